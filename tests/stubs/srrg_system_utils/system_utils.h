// Shape-only stand-in for srrg_core system_utils (tests/stubs/README.md).
#pragma once
#include <string>
namespace srrg_core { double getTime(); }
