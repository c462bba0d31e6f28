// Stand-in for Slam_Utility/src/operate/slam_operations.h: the early-exit macros only.
#ifndef FD_COMPAT_SLAM_OPERATIONS_H_
#define FD_COMPAT_SLAM_OPERATIONS_H_
#define RETURN_FALSE_IF(cond) do { if (cond) { return false; } } while (0)
#define RETURN_TRUE_IF(cond) do { if (cond) { return true; } } while (0)
#define RETURN_FALSE_IF_FALSE(cond) do { if (!(cond)) { return false; } } while (0)
#define RETURN_FALSE_IF_TRUE(cond) do { if (cond) { return false; } } while (0)
#define RETURN_IF(cond) do { if (cond) { return; } } while (0)
#define CONTINUE_IF(cond) if (cond) { continue; }
#define BREAK_IF(cond) if (cond) { break; }
#endif  // FD_COMPAT_SLAM_OPERATIONS_H_
