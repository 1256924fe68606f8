// hyquas_main <file.qasm>: parse -> compile -> run -> print amplitudes + Logger lines.
// Same command line and output contract as the reference driver (main.cpp:233-252 there).
#include "circuit.h"
#include "logger.h"
#include "qasm.h"

int main(int argc, char* argv[]) {
    MyGlobalVars::init();
    if (argc != 2) {
        printf("./parser qasmfile\n");
        exit(1);
    }
    std::unique_ptr<Circuit> c = parse_circuit(std::string(argv[1]));
    c->compile();
    c->run();
    c->printState();
    Logger::print();
    return 0;
}
