// Stand-in for <opencv2/highgui/highgui.hpp> (see core.hpp): nothing on the detect path uses it.
#ifndef JDA_CVSHIM_HIGHGUI_HPP_
#define JDA_CVSHIM_HIGHGUI_HPP_
#include <opencv2/core/core.hpp>
#endif
