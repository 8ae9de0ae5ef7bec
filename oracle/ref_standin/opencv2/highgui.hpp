// opencv2/highgui.hpp stand-in (see ../Eigen/Core): included by FullSystem/CoarseTracker.h, nothing of it is used.
#pragma once
#include "opencv2/imgproc.hpp"
