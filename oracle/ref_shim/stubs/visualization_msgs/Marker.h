// TEST INFRASTRUCTURE: the fields utils.cpp:244-475 (MatrixXd2MarkerArray, off the tracked path) assigns.
#pragma once
#include <string>
namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3 };
    enum { ADD = 0, MODIFY = 0, DELETE = 2 };
    struct { std::string frame_id; } header;
    std::string ns;
    int id = 0, type = 0, action = 0;
    struct { struct { double x = 0, y = 0, z = 0; } position; struct { double x = 0, y = 0, z = 0, w = 1; } orientation; } pose;
    struct { double x = 0, y = 0, z = 0; } scale;
    struct { float r = 0, g = 0, b = 0, a = 0; } color;
};
}  // namespace visualization_msgs
