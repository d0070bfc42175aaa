// Shim (test infrastructure): a small functional stand-in for boost::property_tree::ptree,
// covering exactly the API surface used by the reference's TaskFileParser.cpp and
// LatticeModelFactory.cpp (get / get_optional / get_child / get_child_optional / count / erase / put,
// ordered iteration over (key, subtree) pairs). Paths are '.'-separated like Boost's.
#pragma once
#include <string>
#include <vector>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <utility>
#include <iomanip>
namespace boost {
template <class T> using optional = std::optional<T>;
namespace property_tree {
class ptree_error : public std::runtime_error { public: explicit ptree_error(const std::string &w) : std::runtime_error(w) {} };
class ptree;
template <class Tree> class child_ref {
public:
	child_ref() : _p(nullptr) {}
	explicit child_ref(Tree *p) : _p(p) {}
	explicit operator bool() const { return _p != nullptr; }
	bool operator!() const { return _p == nullptr; }
	Tree &get() const { return *_p; }
	Tree &operator*() const { return *_p; }
	Tree *operator->() const { return _p; }
private:
	Tree *_p;
};
class ptree {
public:
	typedef std::string key_type;
	typedef std::string data_type;
	typedef std::pair<std::string, ptree> value_type;
	typedef std::vector<value_type>::iterator iterator;
	typedef std::vector<value_type>::const_iterator const_iterator;

	ptree() {}
	explicit ptree(const std::string &data) : _data(data) {}

	iterator begin() { return _children.begin(); }
	iterator end() { return _children.end(); }
	const_iterator begin() const { return _children.begin(); }
	const_iterator end() const { return _children.end(); }
	bool empty() const { return _children.empty(); }
	size_t size() const { return _children.size(); }
	std::string &data() { return _data; }
	const std::string &data() const { return _data; }

	size_t count(const std::string &key) const { size_t n = 0; for (auto &c : _children) if (c.first == key) ++n; return n; }
	size_t erase(const std::string &key)
	{
		size_t n = 0;
		for (auto it = _children.begin(); it != _children.end();) { if (it->first == key) { it = _children.erase(it); ++n; } else ++it; }
		return n;
	}
	iterator push_back(const value_type &v) { _children.push_back(v); return _children.end() - 1; }

	child_ref<const ptree> get_child_optional(const std::string &path) const { return child_ref<const ptree>(_walk(path)); }
	child_ref<ptree> get_child_optional(const std::string &path) { return child_ref<ptree>(const_cast<ptree *>(_walk(path))); }
	const ptree &get_child(const std::string &path) const { const ptree *p = _walk(path); if (!p) throw ptree_error("No such node (" + path + ")"); return *p; }
	ptree &get_child(const std::string &path) { ptree *p = const_cast<ptree *>(_walk(path)); if (!p) throw ptree_error("No such node (" + path + ")"); return *p; }

	template <class T> boost::optional<T> get_optional(const std::string &path) const
	{
		const ptree *p = _walk(path);
		if (!p) return boost::optional<T>();
		return _convert<T>(p->_data);
	}
	template <class T> T get(const std::string &path) const
	{
		boost::optional<T> v = get_optional<T>(path);
		if (!v) throw ptree_error("No such node (" + path + ")");
		return *v;
	}
	template <class T> T get_value() const { boost::optional<T> v = _convert<T>(_data); if (!v) throw ptree_error("conversion failed"); return *v; }

	template <class T> ptree &put(const std::string &path, const T &value)
	{
		ptree *node = this;
		size_t pos = 0;
		while (pos <= path.size() && !path.empty())
		{
			size_t dot = path.find('.', pos);
			std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
			ptree *next = nullptr;
			for (auto &c : node->_children) if (c.first == key) { next = &c.second; break; }
			if (!next) { node->_children.push_back(value_type(key, ptree())); next = &node->_children.back().second; }
			node = next;
			if (dot == std::string::npos) break;
			pos = dot + 1;
		}
		std::ostringstream o; o << value; node->_data = o.str();
		return *node;
	}

private:
	const ptree *_walk(const std::string &path) const
	{
		const ptree *node = this;
		if (path.empty()) return node;
		size_t pos = 0;
		while (true)
		{
			size_t dot = path.find('.', pos);
			std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
			const ptree *next = nullptr;
			for (auto &c : node->_children) if (c.first == key) { next = &c.second; break; }
			if (!next) return nullptr;
			node = next;
			if (dot == std::string::npos) break;
			pos = dot + 1;
		}
		return node;
	}
	template <class T> static boost::optional<T> _convert(const std::string &s)
	{
		if constexpr (std::is_same<T, std::string>::value) return s;
		else { std::istringstream i(s); T v; i >> v; if (i.fail()) return boost::optional<T>(); return v; }
	}
	std::string _data;
	std::vector<value_type> _children;
};
}}
