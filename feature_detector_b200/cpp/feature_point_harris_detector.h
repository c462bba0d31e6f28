// Reference include name kept for drop-in builds; the class lives in feature_point_detector.h.
#include "feature_point_detector.h"
