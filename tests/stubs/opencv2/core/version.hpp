// Shape-only stand-in (tests/stubs/README.md): the reference accepts OpenCV 2.x / 3.x (src/types/definitions.h:11-19)
#pragma once
#define CV_MAJOR_VERSION 3
#define CV_MINOR_VERSION 3
#define CV_VERSION_MAJOR 3
