// opencv2/imgproc.hpp stand-in (see ../Eigen/Core): FullSystem/CoarseTracker.h only names cv::Point2f in a declaration.
#pragma once
namespace cv {
struct Point2f { float x, y; };
class Mat;
}  // namespace cv
