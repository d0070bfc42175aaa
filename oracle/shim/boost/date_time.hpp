// Shim (test infrastructure): the slice of boost::posix_time that the reference's
// Log / Timestamp / LoadManager headers touch, on top of <chrono>.
// Lets the UNMODIFIED reference sources under /root/reference/src compile without Boost.
#pragma once
#include <chrono>
#include <ctime>
#include <cstdio>
#include <string>
#include <iomanip>
#include <sstream>
namespace boost { namespace posix_time {
class time_duration {
public:
	time_duration() : _us(0) {}
	time_duration(long h, long m, long s, long frac = 0) : _us(((h * 60 + m) * 60 + s) * 1000000LL + frac) {}
	static time_duration from_us(long long us) { time_duration d; d._us = us; return d; }
	long long total_microseconds() const { return _us; }
	long long total_milliseconds() const { return _us / 1000; }
	long long total_seconds() const { return _us / 1000000; }
	long long _us;
};
class ptime {
public:
	ptime() : _us(0) {}
	explicit ptime(long long us) : _us(us) {}
	long long _us; // microseconds since the unix epoch
};
inline time_duration operator-(const ptime &a, const ptime &b) { return time_duration::from_us(a._us - b._us); }
inline ptime operator+(const ptime &a, const time_duration &d) { return ptime(a._us + d._us); }
inline bool operator<(const ptime &a, const ptime &b) { return a._us < b._us; }
struct microsec_clock {
	static ptime local_time() { return ptime(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count()); }
};
struct second_clock {
	static ptime local_time() { return ptime(std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count() * 1000000LL); }
};
inline std::string to_simple_string(const ptime &t)
{
	std::time_t s = (std::time_t)(t._us / 1000000LL);
	std::tm tm; gmtime_r(&s, &tm);
	static const char *mon[] = { "Jan","Feb","Mar","Apr","May","Jun","Jul","Aug","Sep","Oct","Nov","Dec" };
	char buf[64];
	std::snprintf(buf, sizeof buf, "%04d-%s-%02d %02d:%02d:%02d", tm.tm_year + 1900, mon[tm.tm_mon], tm.tm_mday, tm.tm_hour, tm.tm_min, tm.tm_sec);
	return buf;
}
inline ptime time_from_string(const std::string &str)
{
	int Y = 1970, D = 1, h = 0, m = 0, s = 0; char mo[16] = { 0 };
	static const char *mon[] = { "Jan","Feb","Mar","Apr","May","Jun","Jul","Aug","Sep","Oct","Nov","Dec" };
	int M = 0;
	if (std::sscanf(str.c_str(), "%d-%3[A-Za-z]-%d %d:%d:%d", &Y, mo, &D, &h, &m, &s) >= 3) { for (int i = 0; i < 12; ++i) if (std::string(mo) == mon[i]) M = i; }
	else if (std::sscanf(str.c_str(), "%d-%d-%d %d:%d:%d", &Y, &M, &D, &h, &m, &s) >= 3) M -= 1;
	std::tm tm = {}; tm.tm_year = Y - 1900; tm.tm_mon = M; tm.tm_mday = D; tm.tm_hour = h; tm.tm_min = m; tm.tm_sec = s;
	return ptime((long long)timegm(&tm) * 1000000LL);
}
}}
