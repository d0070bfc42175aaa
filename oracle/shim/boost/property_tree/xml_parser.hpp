// Shim (test infrastructure): read_xml / write_xml for the ptree stand-in. Follows Boost's XML->ptree
// mapping: attributes under "<xmlattr>", text either concatenated into the node data or (flag
// no_concat_text) as "<xmltext>" children, comments as "<xmlcomment>", whitespace-only text dropped.
#pragma once
#include "ptree.hpp"
#include <fstream>
#include <iostream>
#include <cctype>
namespace boost { namespace property_tree { namespace xml_parser {
static const int no_comments = 0x1;
static const int no_concat_text = 0x2;
static const int trim_whitespace = 0x4;
class xml_parser_error : public ptree_error { public: explicit xml_parser_error(const std::string &w) : ptree_error(w) {} };
namespace detail {
inline std::string decode(const std::string &s)
{
	std::string o;
	for (size_t i = 0; i < s.size(); ++i)
	{
		if (s[i] == '&')
		{
			if (s.compare(i, 4, "&lt;") == 0) { o += '<'; i += 3; }
			else if (s.compare(i, 4, "&gt;") == 0) { o += '>'; i += 3; }
			else if (s.compare(i, 5, "&amp;") == 0) { o += '&'; i += 4; }
			else if (s.compare(i, 6, "&quot;") == 0) { o += '"'; i += 5; }
			else if (s.compare(i, 6, "&apos;") == 0) { o += '\''; i += 5; }
			else o += s[i];
		}
		else o += s[i];
	}
	return o;
}
inline std::string encode(const std::string &s)
{
	std::string o;
	for (char c : s) { if (c == '<') o += "&lt;"; else if (c == '>') o += "&gt;"; else if (c == '&') o += "&amp;"; else if (c == '"') o += "&quot;"; else o += c; }
	return o;
}
struct reader {
	const std::string &t; size_t p; int flags;
	reader(const std::string &text, int f) : t(text), p(0), flags(f) {}
	void skipws() { while (p < t.size() && std::isspace((unsigned char)t[p])) ++p; }
	bool starts(const char *s) const { return t.compare(p, std::char_traits<char>::length(s), s) == 0; }
	std::string name() { size_t b = p; while (p < t.size() && !std::isspace((unsigned char)t[p]) && t[p] != '>' && t[p] != '/' && t[p] != '=') ++p; return t.substr(b, p - b); }
	void contents(ptree &node, const std::string &closing)
	{
		while (p < t.size())
		{
			size_t start = p;
			skipws();
			if (p >= t.size()) break;
			if (t[p] == '<')
			{
				if (starts("<!--"))
				{
					size_t e = t.find("-->", p); if (e == std::string::npos) throw xml_parser_error("unterminated comment");
					if (!(flags & no_comments)) node.push_back(ptree::value_type("<xmlcomment>", ptree(t.substr(p + 4, e - p - 4))));
					p = e + 3;
				}
				else if (starts("<?")) { size_t e = t.find("?>", p); if (e == std::string::npos) throw xml_parser_error("unterminated declaration"); p = e + 2; }
				else if (starts("<![CDATA["))
				{
					size_t e = t.find("]]>", p); if (e == std::string::npos) throw xml_parser_error("unterminated cdata");
					text(node, t.substr(p + 9, e - p - 9)); p = e + 3;
				}
				else if (starts("<!")) { size_t e = t.find('>', p); p = e + 1; }
				else if (starts("</"))
				{
					p += 2; std::string n = name(); skipws();
					if (p >= t.size() || t[p] != '>' || n != closing) throw xml_parser_error("mismatched closing tag </" + n + ">");
					++p; return;
				}
				else element(node);
			}
			else
			{
				p = start;
				size_t e = t.find('<', p); if (e == std::string::npos) e = t.size();
				text(node, decode(t.substr(p, e - p))); p = e;
			}
		}
		if (!closing.empty()) throw xml_parser_error("unexpected end of data in <" + closing + ">");
	}
	void text(ptree &node, std::string s)
	{
		if (flags & trim_whitespace)
		{
			size_t b = 0, e = s.size();
			while (b < e && std::isspace((unsigned char)s[b])) ++b;
			while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
			s = s.substr(b, e - b);
		}
		if (flags & no_concat_text) node.push_back(ptree::value_type("<xmltext>", ptree(s)));
		else node.data() += s;
	}
	void element(ptree &parent)
	{
		++p; std::string n = name();
		ptree node;
		ptree attrs; bool hasAttrs = false;
		while (true)
		{
			skipws();
			if (p >= t.size()) throw xml_parser_error("unexpected end of data in tag <" + n + ">");
			if (t[p] == '/') { if (p + 1 < t.size() && t[p + 1] == '>') { p += 2; if (hasAttrs) node.push_back(ptree::value_type("<xmlattr>", attrs)); parent.push_back(ptree::value_type(n, node)); return; } throw xml_parser_error("malformed tag"); }
			if (t[p] == '>') { ++p; break; }
			std::string an = name(); skipws();
			if (p >= t.size() || t[p] != '=') throw xml_parser_error("attribute without value in <" + n + ">");
			++p; skipws();
			char q = t[p]; if (q != '"' && q != '\'') throw xml_parser_error("unquoted attribute value");
			size_t e = t.find(q, p + 1); if (e == std::string::npos) throw xml_parser_error("unterminated attribute value");
			attrs.push_back(ptree::value_type(an, ptree(decode(t.substr(p + 1, e - p - 1))))); hasAttrs = true;
			p = e + 1;
		}
		if (hasAttrs) node.push_back(ptree::value_type("<xmlattr>", attrs));
		contents(node, n);
		parent.push_back(ptree::value_type(n, node));
	}
};
inline void write_node(std::ostream &os, const std::string &key, const ptree &node, int indent)
{
	std::string pad(indent, '\t');
	if (key == "<xmltext>") { os << encode(node.data()); return; }
	if (key == "<xmlcomment>") { os << pad << "<!--" << node.data() << "-->\n"; return; }
	os << pad << "<" << key;
	bool children = false, textOnly = true;
	for (auto &c : node)
	{
		if (c.first == "<xmlattr>") { for (auto &a : c.second) os << " " << a.first << "=\"" << encode(a.second.data()) << "\""; }
		else { children = true; if (c.first != "<xmltext>") textOnly = false; }
	}
	if (!children && node.data().empty()) { os << "/>\n"; return; }
	os << ">";
	if (!node.data().empty()) os << encode(node.data());
	if (children)
	{
		if (!textOnly) os << "\n";
		for (auto &c : node) if (c.first != "<xmlattr>") write_node(os, c.first, c.second, textOnly ? 0 : indent + 1);
		if (!textOnly) os << pad;
	}
	os << "</" << key << ">\n";
}
}
template <class Tree> void read_xml(std::istream &is, Tree &tree, int flags = 0)
{
	std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
	Tree result;
	detail::reader r(text, flags);
	r.contents(result, "");
	tree = result;
}
template <class Tree> void read_xml(const std::string &filename, Tree &tree, int flags = 0)
{
	std::ifstream f(filename.c_str());
	if (!f) throw xml_parser_error("cannot open file " + filename);
	read_xml(f, tree, flags);
}
template <class Tree> void write_xml(std::ostream &os, const Tree &tree)
{
	os << "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n";
	for (auto &c : tree) detail::write_node(os, c.first, c.second, 0);
}
template <class Tree> void write_xml(const std::string &filename, const Tree &tree)
{
	std::ofstream f(filename.c_str());
	if (!f) throw xml_parser_error("cannot open file " + filename);
	write_xml(f, tree);
}
}
using xml_parser::read_xml;
using xml_parser::write_xml;
}}
