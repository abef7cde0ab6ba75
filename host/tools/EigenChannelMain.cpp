// EigenChannelMain.cpp -- command-line entry point: "--config <file>" plus "--name value" overrides,
// like LIA_SpkDet/EigenChannel/src/EigenChannelMain.cpp (channelCompensation JFA).
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config config;
    config.parseCmdLine(argc, argv);
    if (config.existsParam("help")) {
      std::cout << "EigenChannel (lia_ral_b200 engine, eigenChannelMode JFA | LFA): --config <file> [--param value ...]" << std::endl;
      return 0;
    }
    lia::initEngine(config);
    return lia::EigenChannelDispatch(config);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}
