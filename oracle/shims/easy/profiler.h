// oracle/shims -- easy_profiler stand-in: the instrumentation macros expand to nothing.  TEST INFRASTRUCTURE ONLY.
#pragma once
#define EASY_BLOCK(...)
#define EASY_END_BLOCK
#define EASY_FUNCTION(...)
#define EASY_PROFILER_ENABLE
#define EASY_PROFILER_DISABLE
#define EASY_MAIN_THREAD
namespace profiler {
namespace colors {
enum { Red, Green, Blue, Yellow, Orange, Magenta, Cyan, Brown, Black, White, Grey, Purple, Pink, Lime, Amber, Teal, Indigo,
       Navy, Gold, Coral, Olive, DeepOrange, LightBlue, LightGreen, DarkBlue, DarkGreen, DarkRed, DarkTeal, BlueGrey,
       RichRed, RichGreen, RichBlue, RichYellow, Mint, Skin };
}
inline unsigned dumpBlocksToFile(const char*) { return 0; }
}  // namespace profiler
