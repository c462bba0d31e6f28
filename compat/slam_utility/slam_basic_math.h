// Stand-in for Slam_Utility/src/math/slam_basic_math.h: only what the LSD host stage calls.
#ifndef FD_COMPAT_SLAM_BASIC_MATH_H_
#define FD_COMPAT_SLAM_BASIC_MATH_H_
#include "basic_type.h"
namespace Utility {
// GUESS G3: a - b wrapped into (-pi, pi].  Only the host-side LSD region growing uses it.
inline float AngleDiffInRad(float a, float b) {
    float d = a - b;
    while (d > kPai) d -= k2Pai;
    while (d < -kPai) d += k2Pai;
    return d;
}
}  // namespace Utility
#endif  // FD_COMPAT_SLAM_BASIC_MATH_H_
