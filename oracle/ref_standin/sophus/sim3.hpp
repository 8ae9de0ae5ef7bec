// sophus stand-in (see se3.hpp)
#pragma once
#include "sophus/se3.hpp"
