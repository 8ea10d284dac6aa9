// TEST INFRASTRUCTURE: trackdlo.h / utils.h only say `using cv::Mat;` -- the tracker path never touches an image.
#pragma once
namespace cv { class Mat {}; }
