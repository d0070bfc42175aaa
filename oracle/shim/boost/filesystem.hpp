// Shim (test infrastructure): boost::filesystem -> std::filesystem (needs -std=c++17).
#pragma once
#include <filesystem>
#include <fstream>
namespace boost { namespace filesystem = std::filesystem; }
