// Minimal JSON reader for psim model files (objects, arrays, numbers, strings, true/false/null).
// The reference uses nlohmann/json for this (inputManager.cpp:7); the host layer here has no third-party
// dependency.  Throws std::runtime_error with a byte offset on malformed input.
#ifndef PSIM_B200_JSON_H
#define PSIM_B200_JSON_H

#include <cmath>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace psim {

class Json {
public:
    enum class Kind { Null, Bool, Number, String, Array, Object };
    Kind kind = Kind::Null;
    bool boolean = false;
    double number = 0.;
    std::string text;
    std::vector<Json> items;
    std::vector<std::pair<std::string, Json>> members;

    bool contains(const std::string& key) const {
        for (const auto& m : members) {
            if (m.first == key) { return true; }
        }
        return false;
    }
    // same failure mode as nlohmann's at(): missing key is an error the loader reports
    const Json& at(const std::string& key) const {
        if (kind != Kind::Object) { throw std::runtime_error("JSON: '" + key + "' looked up in a non-object"); }
        for (const auto& m : members) {
            if (m.first == key) { return m.second; }
        }
        throw std::runtime_error("JSON: key '" + key + "' not found");
    }
    const Json& at(size_t i) const {
        if (kind != Kind::Array || i >= items.size()) { throw std::runtime_error("JSON: array index out of range"); }
        return items[i];
    }
    size_t size() const { return kind == Kind::Array ? items.size() : members.size(); }
    double num() const {
        if (kind == Kind::Number) { return number; }
        if (kind == Kind::Bool) { return boolean ? 1. : 0.; }
        throw std::runtime_error("JSON: number expected");
    }
    const std::string& str() const {
        if (kind != Kind::String) { throw std::runtime_error("JSON: string expected"); }
        return text;
    }

    static Json parse(const std::string& s) {
        size_t i = 0;
        Json v = value(s, i);
        ws(s, i);
        if (i != s.size()) { fail("trailing characters", i); }
        return v;
    }

private:
    [[noreturn]] static void fail(const char* what, size_t at) {
        throw std::runtime_error(std::string("JSON parse error: ") + what + " at byte " + std::to_string(at));
    }
    static void ws(const std::string& s, size_t& i) {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) { ++i; }
    }
    static std::string string_body(const std::string& s, size_t& i) {
        std::string out;
        ++i;  // opening quote
        while (i < s.size() && s[i] != '"') {
            char c = s[i++];
            if (c == '\\') {
                if (i >= s.size()) { fail("bad escape", i); }
                const char e = s[i++];
                switch (e) {
                case 'n': c = '\n'; break;
                case 't': c = '\t'; break;
                case 'r': c = '\r'; break;
                case 'b': c = '\b'; break;
                case 'f': c = '\f'; break;
                case 'u': {
                    if (i + 4 > s.size()) { fail("bad \\u escape", i); }
                    const unsigned cp = static_cast<unsigned>(std::strtoul(s.substr(i, 4).c_str(), nullptr, 16));
                    i += 4;
                    if (cp < 0x80) {
                        c = static_cast<char>(cp);
                    } else {  // 2- or 3-byte UTF-8 (no surrogate pairs needed for model files)
                        if (cp < 0x800) {
                            out += static_cast<char>(0xC0 | (cp >> 6));
                        } else {
                            out += static_cast<char>(0xE0 | (cp >> 12));
                            out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
                        }
                        c = static_cast<char>(0x80 | (cp & 0x3F));
                    }
                    break;
                }
                default: c = e; break;
                }
            }
            out += c;
        }
        if (i >= s.size()) { fail("unterminated string", i); }
        ++i;
        return out;
    }
    static constexpr int kMaxDepth = 64;  // model files nest four levels deep; '[[[[...' must not overflow the stack
    static Json value(const std::string& s, size_t& i, int depth = 0) {
        if (depth > kMaxDepth) { fail("nesting too deep", i); }
        ws(s, i);
        if (i >= s.size()) { fail("unexpected end", i); }
        Json v;
        const char c = s[i];
        if (c == '{') {
            v.kind = Kind::Object;
            ++i;
            ws(s, i);
            if (i < s.size() && s[i] == '}') {
                ++i;
                return v;
            }
            for (;;) {
                ws(s, i);
                if (i >= s.size() || s[i] != '"') { fail("object key expected", i); }
                std::string key = string_body(s, i);
                ws(s, i);
                if (i >= s.size() || s[i] != ':') { fail("':' expected", i); }
                ++i;
                v.members.emplace_back(std::move(key), value(s, i, depth + 1));
                ws(s, i);
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == '}') {
                    ++i;
                    return v;
                }
                fail("',' or '}' expected", i);
            }
        }
        if (c == '[') {
            v.kind = Kind::Array;
            ++i;
            ws(s, i);
            if (i < s.size() && s[i] == ']') {
                ++i;
                return v;
            }
            for (;;) {
                v.items.push_back(value(s, i, depth + 1));
                ws(s, i);
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == ']') {
                    ++i;
                    return v;
                }
                fail("',' or ']' expected", i);
            }
        }
        if (c == '"') {
            v.kind = Kind::String;
            v.text = string_body(s, i);
            return v;
        }
        if (s.compare(i, 4, "true") == 0) {
            v.kind = Kind::Bool;
            v.boolean = true;
            i += 4;
            return v;
        }
        if (s.compare(i, 5, "false") == 0) {
            v.kind = Kind::Bool;
            i += 5;
            return v;
        }
        if (s.compare(i, 4, "null") == 0) {
            i += 4;
            return v;
        }
        // JSON numbers only: strtod alone would also take "nan", "inf" and hexadecimal floats
        if (!(c == '-' || (c >= '0' && c <= '9'))) { fail("value expected", i); }
        for (size_t k = i; k < s.size() && !(s[k] == ',' || s[k] == '}' || s[k] == ']' || s[k] == ' ' || s[k] == '\n' || s[k] == '\t' || s[k] == '\r'); ++k) {
            const char d = s[k];
            if (!((d >= '0' && d <= '9') || d == '-' || d == '+' || d == '.' || d == 'e' || d == 'E')) { fail("malformed number", k); }
        }
        char* end = nullptr;
        const double d = std::strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) { fail("value expected", i); }
        if (!(d == d) || d > 1.7976931348623157e308 || d < -1.7976931348623157e308) { fail("number out of range", i); }
        v.kind = Kind::Number;
        v.number = d;
        i = static_cast<size_t>(end - s.c_str());
        return v;
    }
};

}  // namespace psim
#endif
