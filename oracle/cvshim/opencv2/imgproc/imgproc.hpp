// Stand-in for <opencv2/imgproc/imgproc.hpp> (see core.hpp).  cv::resize here is nearest-neighbour, NOT OpenCV's
// bilinear: the reference's detect path (fddb.method = 1) always builds the half / quarter images but only samples them
// for nodes with scale != 0, which the shipped model and every model used against this build do not have.
#ifndef JDA_CVSHIM_IMGPROC_HPP_
#define JDA_CVSHIM_IMGPROC_HPP_
#include <opencv2/core/core.hpp>
namespace cv {
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
inline void resize(const Mat &src, Mat &dst, Size dsize, double = 0, double = 0, int = INTER_LINEAR) {
  Mat s = src;  // (dst may alias src)
  Mat d(dsize.height, dsize.width, s.mtype);
  for (int y = 0; y < d.rows; y++)
    for (int x = 0; x < d.cols; x++) {
      const int sy = d.rows ? (int)((long long)y * s.rows / d.rows) : 0, sx = d.cols ? (int)((long long)x * s.cols / d.cols) : 0;
      memcpy(d.data + y * d.step + x * d.elemSize(), s.data + sy * s.step + sx * s.elemSize(), d.elemSize());
    }
  dst = d;
}
}  // namespace cv
#endif
