// oracle/shims: stands in for OpenCV 3.x (the reference's dockerfiles pin ROS kinetic's opencv3 3.3.1; definitions.h:11-19)
#pragma once
#define CV_MAJOR_VERSION 3
#define CV_MINOR_VERSION 3
#define CV_SUBMINOR_VERSION 1
#define CV_VERSION "3.3.1-vslam-shim"
