// Shim (test infrastructure): boost::regex -> std::regex for the unmodified reference sources.
#pragma once
#include <regex>
#include <iomanip>
namespace boost {
using std::regex; using std::smatch; using std::regex_match; using std::regex_search;
}
