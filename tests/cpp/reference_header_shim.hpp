// Stand-in for the reference's trackdlo/include/trackdlo.h AFTER the drop-in recipe of INTEGRATION.md has been applied,
// for boxes where /root/reference does not exist (tests/test_dropin.py uses the real, patched header when it does).
// It reproduces what matters about that file's structure: third-party includes first (trackdlo.h:3-45; satisfied by
// oracle/ref_shim: the Eigen stand-in + empty ROS / OpenCV / PCL headers), then the include guard that is ALREADY OPEN
// (trackdlo.h:47-48) and the two using-declarations (:50-51) when the adapter is included where the class declaration
// used to be (:53-130), then the guard's #endif (:131).
#pragma once

#include <Eigen/Dense>
#include <vector>
#include <ros/ros.h>
#include <opencv2/core/core.hpp>
#include <visualization_msgs/MarkerArray.h>
#include <string>

#ifndef TRACKDLO_H
#define TRACKDLO_H

using Eigen::MatrixXd;
using cv::Mat;

#include <trackdlo_adapter.hpp>      // <- the recipe: replaces the class declaration

#endif
