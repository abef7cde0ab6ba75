// EstimateDMatrixMain.cpp -- command-line entry point: "--config <file>" plus "--name value" overrides,
// like LIA_SpkDet/EstimateDMatrix/src/EstimateDMatrixMain.cpp.
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config config;
    config.parseCmdLine(argc, argv);
    if (config.existsParam("help")) {
      std::cout << "EstimateDMatrix (lia_ral_b200 engine): --config <file> [--param value ...]" << std::endl;
      return 0;
    }
    lia::initEngine(config);
    return lia::EstimateDMatrix(config);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}
