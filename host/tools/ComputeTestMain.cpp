// ComputeTestMain.cpp -- command-line entry point: "--config <file>" plus "--name value" overrides,
// like LIA_SpkDet/ComputeTest/src/ComputeTestMain.cpp.
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config config;
    config.parseCmdLine(argc, argv);
    if (config.existsParam("help")) {
      std::cout << "ComputeTest (lia_ral_b200 engine): --config <file> [--param value ...]" << std::endl;
      return 0;
    }
    lia::initEngine(config);  // device + (several ranks) the NCCL communicator
    return lia::ComputeTestDispatch(config);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}
