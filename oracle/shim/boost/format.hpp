// Shim (test infrastructure): minimal boost::format (printf-style "%d %f %s" directives fed with operator%),
// enough for the reference's .ldf debug writer (LatticeModelFactory.cpp:1000-1034).
#pragma once
#include <string>
#include <vector>
#include <sstream>
#include <fstream>
#include <cstdio>
namespace boost {
class format {
public:
	explicit format(const std::string &fmt) : _fmt(fmt) {}
	template <class T> format &operator%(const T &v) { std::ostringstream o; o << std::fixed; o.precision(6); o << v; _args.push_back(o.str()); return *this; }
	std::string str() const
	{
		std::string out; size_t arg = 0;
		for (size_t i = 0; i < _fmt.size(); ++i)
		{
			if (_fmt[i] == '%' && i + 1 < _fmt.size())
			{
				if (_fmt[i + 1] == '%') { out += '%'; ++i; continue; }
				size_t j = i + 1;
				while (j < _fmt.size() && !std::isalpha((unsigned char)_fmt[j])) ++j;
				if (arg < _args.size()) out += _args[arg++];
				i = j;
			}
			else out += _fmt[i];
		}
		return out;
	}
private:
	std::string _fmt; std::vector<std::string> _args;
};
inline std::ostream &operator<<(std::ostream &os, const format &f) { return os << f.str(); }
}
