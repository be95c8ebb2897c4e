// json_reader.h — the JSON reader of the native scene loader (scene_io.cpp: the header of a .crts file, a glTF document), host
// code only. A small DOM that keeps what matters of the behaviour of the libraries the reference reads these files with.
#pragma once

#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace crt_json {

// The reference reads the header of a .crts file with nlohmann::json (util/scene.cpp:417-425); what matters of that
// library's behaviour is restated here: numbers without fraction or exponent are integers (unsigned if not negative), the
// others go through strtod; get<float>() is a static_cast from whichever of the three the number is; a key given twice
// keeps its last value; operator[] on a missing key yields null, whose size() is 0.
struct Json {
    enum Kind { kNull, kBool, kUnsigned, kSigned, kFloat, kString, kArray, kObject };
    Kind kind = kNull;
    bool boolean = false;
    uint64_t u = 0;
    int64_t i = 0;
    double d = 0;
    std::string str;
    std::vector<Json> items;
    std::vector<std::pair<std::string, Json>> members;

    const Json *find(const char *key) const
    {
        const Json *found = nullptr;
        for (const auto &m : members) {
            if (m.first == key) {
                found = &m.second;
            }
        }
        return found;
    }
    const Json &at(const char *key, const std::string &where) const
    {
        const Json *j = kind == kObject ? find(key) : nullptr;
        if (!j) {
            throw std::runtime_error("scene header: " + where + " has no \"" + key + "\"");
        }
        return *j;
    }
    const Json &at(size_t index, const std::string &where) const
    {
        if (kind != kArray || index >= items.size()) {
            throw std::runtime_error("scene header: " + where + " has no element " + std::to_string(index));
        }
        return items[index];
    }
    size_t size() const
    {
        return kind == kArray ? items.size() : (kind == kObject ? members.size() : (kind == kNull ? 0 : 1));
    }
    template <typename T>
    T number(const std::string &where) const
    {
        switch (kind) {
        case kUnsigned: return static_cast<T>(u);
        case kSigned: return static_cast<T>(i);
        case kFloat: return static_cast<T>(d);
        case kBool: return static_cast<T>(boolean);
        default: throw std::runtime_error("scene header: " + where + " is not a number");
        }
    }
    const std::string &string(const std::string &where) const
    {
        if (kind != kString) {
            throw std::runtime_error("scene header: " + where + " is not a string");
        }
        return str;
    }
    std::vector<float> floats(size_t at_least, const std::string &where) const
    {
        if (kind != kArray || items.size() < at_least) {
            throw std::runtime_error("scene header: " + where + " is not an array of " + std::to_string(at_least) + " numbers");
        }
        std::vector<float> out(items.size());
        for (size_t k = 0; k < items.size(); ++k) {
            out[k] = items[k].number<float>(where);
        }
        return out;
    }
};

class JsonParser {
public:
    JsonParser(const char *begin, const char *end) : p(begin), end(end) {}
    Json parse_document()
    {
        Json j = value(0);
        skip_space();
        if (p != end) {
            fail("text after the document");
        }
        return j;
    }

private:
    const char *p, *end;
    [[noreturn]] void fail(const std::string &what) const
    {
        throw std::runtime_error("scene header: malformed JSON (" + what + ")");
    }
    void skip_space()
    {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) {
            ++p;
        }
    }
    bool literal(const char *word)
    {
        const size_t n = std::strlen(word);
        if ((size_t)(end - p) >= n && std::memcmp(p, word, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    static void append_utf8(std::string &out, uint32_t cp)
    {
        if (cp < 0x80) {
            out += (char)cp;
        } else if (cp < 0x800) {
            out += (char)(0xC0 | (cp >> 6));
            out += (char)(0x80 | (cp & 0x3F));
        } else if (cp < 0x10000) {
            out += (char)(0xE0 | (cp >> 12));
            out += (char)(0x80 | ((cp >> 6) & 0x3F));
            out += (char)(0x80 | (cp & 0x3F));
        } else {
            out += (char)(0xF0 | (cp >> 18));
            out += (char)(0x80 | ((cp >> 12) & 0x3F));
            out += (char)(0x80 | ((cp >> 6) & 0x3F));
            out += (char)(0x80 | (cp & 0x3F));
        }
    }
    uint32_t hex4()
    {
        if (end - p < 4) {
            fail("short \\u escape");
        }
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k, ++p) {
            const char c = *p;
            v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : (fail("bad \\u escape"), 0));
        }
        return v;
    }
    std::string string_body()
    {
        std::string out;
        ++p;  // the opening quote
        for (;;) {
            if (p >= end) {
                fail("unterminated string");
            }
            const char c = *p++;
            if (c == '"') {
                return out;
            }
            if (c != '\\') {
                out += c;
                continue;
            }
            if (p >= end) {
                fail("unterminated escape");
            }
            const char e = *p++;
            switch (e) {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u': {
                uint32_t cp = hex4();
                if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                    p += 2;
                    const uint32_t low = hex4();
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (low - 0xDC00);
                }
                append_utf8(out, cp);
                break;
            }
            default: fail("unknown escape");
            }
        }
    }
    Json number()
    {
        const char *start = p;
        bool integral = true;
        if (p < end && *p == '-') {
            ++p;
        }
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
            integral = integral && *p >= '0' && *p <= '9';
            ++p;
        }
        const std::string text(start, p);
        if (text.empty() || text == "-") {
            fail("a value was expected");
        }
        Json j;
        char *stop = nullptr;
        errno = 0;
        if (integral && text[0] != '-') {
            j.u = std::strtoull(text.c_str(), &stop, 10);
            j.kind = Json::kUnsigned;
        } else if (integral) {
            j.i = std::strtoll(text.c_str(), &stop, 10);
            j.kind = Json::kSigned;
        }
        if (!integral || errno == ERANGE) {  // (an integer too large for 64 bits is read as a float, as nlohmann does)
            j.d = std::strtod(text.c_str(), &stop);
            j.kind = Json::kFloat;
        }
        if (!stop || *stop != '\0') {
            fail("bad number " + text);
        }
        return j;
    }
    Json value(int depth)
    {
        if (depth > 64) {
            fail("nesting too deep");
        }
        skip_space();
        if (p >= end) {
            fail("unexpected end");
        }
        Json j;
        if (*p == '{') {
            ++p;
            j.kind = Json::kObject;
            skip_space();
            if (p < end && *p == '}') {
                ++p;
                return j;
            }
            for (;;) {
                skip_space();
                if (p >= end || *p != '"') {
                    fail("a member name was expected");
                }
                std::string key = string_body();
                skip_space();
                if (p >= end || *p != ':') {
                    fail("':' was expected");
                }
                ++p;
                j.members.emplace_back(std::move(key), value(depth + 1));
                skip_space();
                if (p < end && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < end && *p == '}') {
                    ++p;
                    return j;
                }
                fail("',' or '}' was expected");
            }
        }
        if (*p == '[') {
            ++p;
            j.kind = Json::kArray;
            skip_space();
            if (p < end && *p == ']') {
                ++p;
                return j;
            }
            for (;;) {
                j.items.push_back(value(depth + 1));
                skip_space();
                if (p < end && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < end && *p == ']') {
                    ++p;
                    return j;
                }
                fail("',' or ']' was expected");
            }
        }
        if (*p == '"') {
            j.kind = Json::kString;
            j.str = string_body();
            return j;
        }
        if (literal("true")) {
            j.kind = Json::kBool;
            j.boolean = true;
            return j;
        }
        if (literal("false")) {
            j.kind = Json::kBool;
            return j;
        }
        if (literal("null")) {
            return j;
        }
        return number();
    }
};

}  // namespace crt_json
