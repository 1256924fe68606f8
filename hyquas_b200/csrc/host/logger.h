// printf-style line buffer flushed as "Logger[rank]: ..." -- the same observable contract as the
// reference's Logger (src/logger.h:9-49): harnesses grep for "Logger" / "Time Cost".
#pragma once
#include <cstdarg>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "utils.h"

class Logger {
public:
    static void add(const char* format, ...) {
        char line[1024];
        va_list ap;
        va_start(ap, format);
        vsnprintf(line, sizeof(line), format, ap);
        va_end(ap);
        lines().emplace_back(line);
    }
    static void print() {
        std::string who = MyGlobalVars::numGPUs > 1 ? "[" + std::to_string(MyMPI::rank) + "]" : "";
        for (const auto& s : lines()) std::cout << "Logger" << who << ": " << s << std::endl;
        lines().clear();
    }
    static const std::vector<std::string>& pending() { return lines(); }
private:
    static std::vector<std::string>& lines() {
        static std::vector<std::string> buf;
        return buf;
    }
};
