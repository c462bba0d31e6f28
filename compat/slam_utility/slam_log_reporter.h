// Stand-in for Slam_Utility/src/log/slam_log_reporter.h: stream macros; the hot path logs nothing.
#ifndef FD_COMPAT_SLAM_LOG_REPORTER_H_
#define FD_COMPAT_SLAM_LOG_REPORTER_H_
#include <iostream>
#define ReportInfo(...) do { std::cout << __VA_ARGS__ << std::endl; } while (0)
#define ReportError(...) do { std::cerr << __VA_ARGS__ << std::endl; } while (0)
#define ReportWarn(...) do { std::cout << __VA_ARGS__ << std::endl; } while (0)
#define ReportColorWarn(...) do { std::cout << __VA_ARGS__ << std::endl; } while (0)
#define ReportText(...) do { std::cout << __VA_ARGS__; } while (0)
#endif  // FD_COMPAT_SLAM_LOG_REPORTER_H_
