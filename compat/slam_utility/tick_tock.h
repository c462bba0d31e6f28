// Stand-in for Slam_Utility/src/tick_tock/tick_tock.h: nothing of it is used by the sources compiled here.
#ifndef FD_COMPAT_TICK_TOCK_H_
#define FD_COMPAT_TICK_TOCK_H_
#endif  // FD_COMPAT_TICK_TOCK_H_
