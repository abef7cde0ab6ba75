// TrainTargetMain.cpp -- command-line entry point: "--config <file>" plus "--name value" overrides,
// like LIA_SpkDet/TrainTarget/src/TrainTargetMain.cpp.
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config config;
    config.parseCmdLine(argc, argv);
    if (config.existsParam("help")) {
      std::cout << "TrainTarget (lia_ral_b200 engine): --config <file> [--param value ...]" << std::endl;
      return 0;
    }
    lia::initEngine(config);  // device + (several ranks) the NCCL communicator
    return lia::TrainTargetDispatch(config);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}
