// ComputeJFAStatsMain.cpp -- command-line entry point: "--config <file>" plus "--name value" overrides,
// like LIA_SpkDet/ComputeJFAStats/src/ComputeJFAStatsMain.cpp.
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config config;
    config.parseCmdLine(argc, argv);
    if (config.existsParam("help")) {
      std::cout << "ComputeJFAStats (lia_ral_b200 engine): --config <file> [--param value ...]" << std::endl;
      return 0;
    }
    lia::initEngine(config);
    return lia::ComputeJFAStats(config);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}
