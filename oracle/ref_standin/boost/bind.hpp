// boost/bind.hpp stand-in (see ../Eigen/Core): util/settings.cpp includes it without using it.
#pragma once
