// OpenQASM-2 subset front end: the input dialect of the reference driver (main.cpp:29-231 there):
// statements `cx ccx cy cz h x y z s sdg t tdg` and `crx cry crz cu1 u1 u3 rx ry rz` with parameters written
// `pi*x`, `pi/x` or a plain decimal; operands are one token without spaces; `//`, OPENQASM, include lines skipped.
#pragma once
#include <memory>
#include <string>

#include "circuit.h"

namespace hyquas {
// On malformed input: message into `err` and nullptr (the CLI prints it and exits 1 like the reference).
std::unique_ptr<Circuit> parseQasmText(const std::string& text, std::string& err);
std::unique_ptr<Circuit> parseQasmFile(const std::string& filename, std::string& err);
}

// reference-compatible entry point (prints + exit(1) on failure)
std::unique_ptr<Circuit> parse_circuit(const std::string& filename);
