// TEST INFRASTRUCTURE ONLY — shadows the reference's FullSystem/FullSystem.h when its hot-path translation units are compiled
// into oracle/_ref/: the real header drags in IMU/GTSAM, yaml-cpp, the pixel selector and the initialiser, none of which the
// compiled functions (PointFrameResidual::linearize/applyRes, EFResidual::takeDataF, EnergyFunctional::*) touch.  It forwards to
// the reference's own headers for everything those functions do use.
#pragma once
#include "util/NumType.h"
#include "util/globalCalib.h"
#include "vector"
#include <iostream>
#include <fstream>
#include <deque>
#include <math.h>
#include "FullSystem/Residuals.h"
#include "FullSystem/HessianBlocks.h"
#include "util/FrameShell.h"
#include "util/IndexThreadReduce.h"
#include "OptimizationBackend/EnergyFunctional.h"
namespace dso {
class FullSystem;
}
namespace dmvio {
// The GTSAM bridge is only reached when setting_useGTSAMIntegration is true (OptimizationBackend/EnergyFunctional.cpp:L335, L545, L958);
// the harness runs the reference's own no-IMU branch (L971-973), so these bodies are never executed.
class BAGTSAMIntegration {
 public:
  static void unreachable() { fprintf(stderr, "oracle/_ref: BAGTSAMIntegration reached (setting_useGTSAMIntegration must be false)\n"); abort(); }
  void updateBAValues(std::vector<dso::EFFrame*>&) { unreachable(); }
  double getBAEnergy(bool) { unreachable(); return 0; }
  void addMarginalizedPointsBA(const dso::MatXX&, const dso::VecX&, std::vector<dso::EFFrame*>&) { unreachable(); }
  void addPriorBA(dso::EFFrame*, const dso::Vec8&, const dso::Vec8&) { unreachable(); }
  void marginalizeBAFrame(dso::EFFrame*) { unreachable(); }
  dso::VecX computeBAUpdate(const dso::MatXX&, const dso::VecX&, double, std::vector<dso::EFFrame*>&, const dso::MatXX&) { unreachable(); return dso::VecX(); }
};
}  // namespace dmvio
