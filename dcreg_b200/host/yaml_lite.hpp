// yaml_lite.hpp - the subset of YAML that DCReg's config files use (DCReg/config/*.yaml), dependency-free.
//
// The reference parses its configs with yaml-cpp (DCReg/src/icp_test_runner.cpp:20-153); yaml-cpp is not available
// here, and the configs only use: nested block mappings by indentation, scalar values (numbers, booleans, bare or
// quoted strings), flow sequences of scalars (`[ "A", "B" ]`), quoted keys, `#` comments and blank lines.
#pragma once
#include <cctype>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace yaml_lite {

struct Node {
    bool defined = false;
    std::string scalar;                         // scalar value (unquoted)
    std::vector<std::string> seq;               // flow sequence items
    std::vector<std::pair<std::string, std::shared_ptr<Node>>> map;   // insertion order kept

    explicit operator bool() const { return defined; }
    const Node& operator[](const std::string& key) const {
        static const Node none;
        for (const auto& kv : map)
            if (kv.first == key) return *kv.second;
        return none;
    }
    template <typename T> T as() const;
};

struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };

inline std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

inline std::string strip_comment(const std::string& line) {
    bool in_s = false, in_d = false;
    for (size_t i = 0; i < line.size(); ++i) {
        const char c = line[i];
        if (c == '\'' && !in_d) in_s = !in_s;
        else if (c == '"' && !in_s) in_d = !in_d;
        else if (c == '#' && !in_s && !in_d && (i == 0 || std::isspace((unsigned char)line[i - 1]))) return line.substr(0, i);
    }
    return line;
}

inline std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
        return s.substr(1, s.size() - 2);
    return s;
}

inline std::vector<std::string> parse_flow_seq(const std::string& v) {
    std::vector<std::string> out;
    std::string cur;
    bool in_s = false, in_d = false;
    for (size_t i = 1; i + 1 < v.size(); ++i) {
        const char c = v[i];
        if (c == '\'' && !in_d) in_s = !in_s;
        if (c == '"' && !in_s) in_d = !in_d;
        if (c == ',' && !in_s && !in_d) { out.push_back(unquote(trim(cur))); cur.clear(); }
        else cur.push_back(c);
    }
    if (!trim(cur).empty()) out.push_back(unquote(trim(cur)));
    return out;
}

template <> inline std::string Node::as<std::string>() const {
    if (!defined) throw ParseError("missing key");
    return scalar;
}
template <> inline double Node::as<double>() const {
    if (!defined) throw ParseError("missing key");
    size_t pos = 0;
    const double v = std::stod(scalar, &pos);
    if (pos != scalar.size()) throw ParseError("bad number: " + scalar);
    return v;
}
template <> inline int Node::as<int>() const {
    if (!defined) throw ParseError("missing key");
    size_t pos = 0;
    const long v = std::stol(scalar, &pos);
    if (pos != scalar.size()) throw ParseError("bad integer: " + scalar);
    return (int)v;
}
template <> inline bool Node::as<bool>() const {
    if (!defined) throw ParseError("missing key");
    std::string s;
    for (char c : scalar) s.push_back((char)std::tolower((unsigned char)c));
    if (s == "true" || s == "yes" || s == "on") return true;
    if (s == "false" || s == "no" || s == "off") return false;
    throw ParseError("bad boolean: " + scalar);
}
template <> inline std::vector<std::string> Node::as<std::vector<std::string>>() const {
    if (!defined) throw ParseError("missing key");
    return seq;
}

inline Node parse(std::istream& in) {
    Node root;
    root.defined = true;
    std::vector<std::pair<int, Node*>> stack;   // (indent of the keys inside this mapping, node)
    stack.push_back({-1, &root});
    std::string raw;
    int lineno = 0;
    while (std::getline(in, raw)) {
        ++lineno;
        if (!raw.empty() && raw.back() == '\r') raw.pop_back();
        const std::string line = strip_comment(raw);
        if (trim(line).empty()) continue;
        int indent = 0;
        while (indent < (int)line.size() && line[indent] == ' ') ++indent;
        if (indent < (int)line.size() && line[indent] == '\t') throw ParseError("tab indentation at line " + std::to_string(lineno));
        const std::string body = trim(line);
        // key: find the first ':' outside quotes that is followed by space or end of line
        size_t colon = std::string::npos;
        bool in_s = false, in_d = false;
        for (size_t i = 0; i < body.size(); ++i) {
            const char c = body[i];
            if (c == '\'' && !in_d) in_s = !in_s;
            else if (c == '"' && !in_s) in_d = !in_d;
            else if (c == ':' && !in_s && !in_d && (i + 1 == body.size() || body[i + 1] == ' ')) { colon = i; break; }
        }
        if (colon == std::string::npos) throw ParseError("expected 'key: value' at line " + std::to_string(lineno));
        const std::string key = unquote(trim(body.substr(0, colon)));
        const std::string val = trim(body.substr(colon + 1));
        while (stack.size() > 1 && indent <= stack.back().first) stack.pop_back();
        Node* parent = stack.back().second;
        auto child = std::make_shared<Node>();
        child->defined = true;
        if (val.empty()) {
            stack.push_back({indent, child.get()});             // nested mapping follows
        } else if (val.front() == '[') {
            if (val.back() != ']') throw ParseError("unterminated flow sequence at line " + std::to_string(lineno));
            child->seq = parse_flow_seq(val);
        } else {
            child->scalar = unquote(val);
        }
        bool replaced = false;
        for (auto& kv : parent->map)
            if (kv.first == key) { kv.second = child; replaced = true; }
        if (!replaced) parent->map.push_back({key, child});
    }
    return root;
}

inline Node load_file(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) throw ParseError("cannot open " + path);
    return parse(f);
}

}  // namespace yaml_lite
