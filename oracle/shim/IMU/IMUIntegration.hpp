// TEST INFRASTRUCTURE ONLY — shadows the reference's IMU/IMUIntegration.hpp (GTSAM-based) for the oracle/_ref build.
// CoarseTracker only reaches these calls when dso::setting_useIMU is true (FullSystem/CoarseTracker.cpp:L612, L708, L765);
// the harness sets it to false, i.e. runs the reference's own visual-only branch (L639-683).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <array>
#include "util/NumType.h"
#include "util/IndexThreadReduce.h"  // FullSystem/CoarseInitializer.h names IndexThreadReduce / std::array after including this header
namespace dmvio {
class IMUIntegration {
 public:
  static void unreachable() { fprintf(stderr, "oracle/_ref: IMUIntegration reached (setting_useIMU must be false)\n"); abort(); }
  bool isCoarseInitialized() { unreachable(); return false; }
  dso::SE3 computeCoarseUpdate(const dso::Mat88&, const dso::Vec8&, float, float, double&, double&, double&) { unreachable(); return dso::SE3(); }
  void acceptCoarseUpdate() { unreachable(); }
  void addVisualToCoarseGraph(const dso::Mat88&, const dso::Vec8&, bool) { unreachable(); }
};
}  // namespace dmvio
