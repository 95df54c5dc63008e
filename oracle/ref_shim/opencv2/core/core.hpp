// Stand-in for OpenCV's core header: cv::Mat only carries the debug image pointer the reference's
// getDebugDisplay returns (PI/costs.cu:271-284).  TEST INFRASTRUCTURE (oracle/refbuild.py).
#ifndef REF_SHIM_OPENCV_CORE_
#define REF_SHIM_OPENCV_CORE_
#define CV_32F 5
namespace cv {
class Mat {
 public:
  Mat() : rows(0), cols(0), type_(0), data(nullptr) {}
  Mat(int r, int c, int type, void *d) : rows(r), cols(c), type_(type), data(d) {}
  int rows, cols, type_;
  void *data;
};
}  // namespace cv
#endif
