#include "qasm.h"

#include <cmath>
#include <fstream>
#include <sstream>
#include <vector>

namespace hyquas {
namespace {

// every maximal run of digits inside the operand token is a qubit index
std::vector<int> qubitIds(const std::string& tok) {
    std::vector<int> ids;
    size_t i = 0;
    while (i < tok.size()) {
        if (isdigit((unsigned char)tok[i])) {
            int v = 0;
            while (i < tok.size() && isdigit((unsigned char)tok[i])) v = v * 10 + (tok[i++] - '0');
            ids.push_back(v);
        } else {
            i++;
        }
    }
    return ids;
}

bool parseAngle(const std::string& s, qreal& out) {
    const qreal pi = acos(-1);
    try {
        if (s.compare(0, 3, "pi*") == 0) out = pi * std::stod(s.substr(3));
        else if (s.compare(0, 3, "pi/") == 0) out = pi / std::stod(s.substr(3));
        else out = 1.0 * std::stod(s);
    } catch (...) {
        return false;
    }
    return true;
}

struct Spec { const char* name; int params; int qubits; };
const Spec kSpecs[] = {
    {"cx", 0, 2}, {"ccx", 0, 3}, {"cy", 0, 2}, {"cz", 0, 2}, {"h", 0, 1}, {"x", 0, 1}, {"y", 0, 1}, {"z", 0, 1},
    {"s", 0, 1}, {"sdg", 0, 1}, {"t", 0, 1}, {"tdg", 0, 1},
    {"crx", 1, 2}, {"cry", 1, 2}, {"crz", 1, 2}, {"cu1", 1, 2}, {"u1", 1, 1}, {"u3", 3, 1}, {"rx", 1, 1}, {"ry", 1, 1}, {"rz", 1, 1},
};

Gate makeGate(const std::string& n, const std::vector<int>& q, const std::vector<qreal>& p) {
    if (n == "cx") return Gate::CNOT(q[0], q[1]);
    if (n == "ccx") return Gate::CCX(q[0], q[1], q[2]);
    if (n == "cy") return Gate::CY(q[0], q[1]);
    if (n == "cz") return Gate::CZ(q[0], q[1]);
    if (n == "h") return Gate::H(q[0]);
    if (n == "x") return Gate::X(q[0]);
    if (n == "y") return Gate::Y(q[0]);
    if (n == "z") return Gate::Z(q[0]);
    if (n == "s") return Gate::S(q[0]);
    if (n == "sdg") return Gate::SDG(q[0]);
    if (n == "t") return Gate::T(q[0]);
    if (n == "tdg") return Gate::TDG(q[0]);
    if (n == "crx") return Gate::CRX(q[0], q[1], p[0]);
    if (n == "cry") return Gate::CRY(q[0], q[1], p[0]);
    if (n == "crz") return Gate::CRZ(q[0], q[1], p[0]);
    if (n == "cu1") return Gate::CU1(q[0], q[1], p[0]);
    if (n == "u1") return Gate::U1(q[0], p[0]);
    if (n == "u3") return Gate::U3(q[0], p[0], p[1], p[2]);
    if (n == "rx") return Gate::RX(q[0], p[0]);
    if (n == "ry") return Gate::RY(q[0], p[0]);
    return Gate::RZ(q[0], p[0]);
}

}  // namespace

std::unique_ptr<Circuit> parseQasmText(const std::string& text, std::string& err) {
    std::unique_ptr<Circuit> c;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string head, operand;
        if (!(ls >> head)) continue;
        if (head == "//" || head == "OPENQASM" || head == "include") continue;
        if (head == "qreg") {
            ls >> operand;
            std::vector<int> ids = qubitIds(operand);
            if (ids.empty()) { err = "fail to load circuit"; return nullptr; }
            c.reset(new Circuit(ids[0]));
            continue;
        }
        std::string name = head.substr(0, head.find('('));
        const Spec* spec = nullptr;
        for (const Spec& s : kSpecs) if (name == s.name) spec = &s;
        if (!spec || (spec->params == 0) != (head.find('(') == std::string::npos)) {
            err = "unrecognized token " + head;
            return nullptr;
        }
        std::vector<qreal> params;
        if (spec->params) {
            const size_t l = head.find('('), r = head.rfind(')');
            if (r == std::string::npos || r < l) { err = "unrecognized token " + head; return nullptr; }
            std::stringstream ps(head.substr(l + 1, r - l - 1));
            std::string item;
            while (std::getline(ps, item, ',')) {
                qreal v;
                if (!parseAngle(item, v)) { err = "bad parameter in " + head; return nullptr; }
                params.push_back(v);
            }
        }
        ls >> operand;
        std::vector<int> q = qubitIds(operand);
        if (!c) { err = "gate before qreg"; return nullptr; }
        if ((int)params.size() != spec->params || (int)q.size() != spec->qubits) { err = "wrong operand count for " + head; return nullptr; }
        for (int id : q) if (id < 0 || id >= c->numQubits) { err = "qubit index out of range in " + line; return nullptr; }
        for (size_t a = 0; a < q.size(); a++)
            for (size_t b = a + 1; b < q.size(); b++)
                if (q[a] == q[b]) { err = "repeated qubit in " + line; return nullptr; }
        c->addGate(makeGate(name, q, params));
    }
    if (!c) err = "fail to load circuit";
    return c;
}

std::unique_ptr<Circuit> parseQasmFile(const std::string& filename, std::string& err) {
    std::ifstream f(filename);
    if (!f) { err = "fail to open " + filename; return nullptr; }
    std::stringstream ss;
    ss << f.rdbuf();
    return parseQasmText(ss.str(), err);
}

}  // namespace hyquas

std::unique_ptr<Circuit> parse_circuit(const std::string& filename) {
    std::string err;
    auto c = hyquas::parseQasmFile(filename, err);
    if (!c) {
        printf("%s\n", err.c_str());
        exit(1);
    }
    return c;
}
