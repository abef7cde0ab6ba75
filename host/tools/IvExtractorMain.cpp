// IvExtractorMain.cpp -- command-line entry point: "--config <file>" plus "--name value" overrides,
// like LIA_SpkDet/IvExtractor/src/IvExtractorMain.cpp.
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config config;
    config.parseCmdLine(argc, argv);
    if (config.existsParam("help")) {
      std::cout << "IvExtractor (lia_ral_b200 engine): --config <file> [--param value ...]" << std::endl;
      return 0;
    }
    lia::initEngine(config);  // device + (several ranks) the NCCL communicator
    // IvExtractorMain.cpp:99-111: classic (default) | ubmWeight | eigenDecomposition
    const std::string mode = config.existsParam("mode") ? config.getParam("mode") : "classic";
    if (mode == "classic") return lia::IvExtractor(config);
    if (mode == "ubmWeight") return lia::IvExtractorUbmWeigth(config);
    if (mode == "eigenDecomposition") return lia::IvExtractorEigenDecomposition(config);
    std::cout << "Wrong mode, should be classic | ubmWeight | eigenDecomposition" << std::endl;
    return 0;
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}
