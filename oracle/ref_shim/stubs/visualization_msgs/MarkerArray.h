// TEST INFRASTRUCTURE: see Marker.h.
#pragma once
#include <vector>
#include "Marker.h"
namespace visualization_msgs { struct MarkerArray { std::vector<Marker> markers; }; }
