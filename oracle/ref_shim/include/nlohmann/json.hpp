// nlohmann/json.hpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim).
//
// Stand-in for the nlohmann::json surface general.cpp:19-51 uses: stream extraction of one flat
// object, operator[](key) (a missing key yields a null value), empty() (true for null) and
// implicit conversion to std::string / int / double.
#pragma once
#include <cassert>
#include <cstdlib>
#include <fstream>
#include <istream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>

namespace nlohmann {

class json {
public:
    enum class kind { null, string, number, boolean, object };
    json() : k_(kind::null) {}
    json &operator[](const std::string &key) {
        if (k_ == kind::null) k_ = kind::object;
        if (k_ != kind::object) throw std::domain_error("json shim: operator[] on a non-object");
        return obj_[key];
    }
    bool empty() const { return k_ == kind::null || (k_ == kind::object && obj_.empty()); }

    template<typename T, typename std::enable_if<std::is_same<T, std::string>::value ||
                                                         (std::is_arithmetic<T>::value && !std::is_same<T, char>::value),
                                                 int>::type = 0>
    operator T() const {
        return get<T>();
    }

    friend std::istream &operator>>(std::istream &is, json &j) {
        std::stringstream ss;
        ss << is.rdbuf();
        const std::string text = ss.str();
        size_t p = 0;
        j = parse_value(text, p);
        return is;
    }

private:
    template<typename T>
    typename std::enable_if<std::is_same<T, std::string>::value, T>::type get() const {
        if (k_ != kind::string) throw std::domain_error("json shim: value is not a string");
        return str_;
    }
    template<typename T>
    typename std::enable_if<std::is_arithmetic<T>::value, T>::type get() const {
        if (k_ != kind::number && k_ != kind::boolean) throw std::domain_error("json shim: value is not a number");
        return static_cast<T>(num_);
    }
    static void skip(const std::string &t, size_t &p) {
        while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) ++p;
    }
    static std::string parse_string(const std::string &t, size_t &p) {
        if (t[p] != '"') throw std::invalid_argument("json shim: expected string");
        std::string s;
        for (++p; p < t.size() && t[p] != '"'; ++p) {
            if (t[p] == '\\' && p + 1 < t.size()) ++p;
            s.push_back(t[p]);
        }
        if (p >= t.size()) throw std::invalid_argument("json shim: unterminated string");
        ++p;
        return s;
    }
    static json parse_value(const std::string &t, size_t &p) {
        skip(t, p);
        if (p >= t.size()) throw std::invalid_argument("json shim: unexpected end of input");
        json j;
        if (t[p] == '{') {
            j.k_ = kind::object;
            ++p;
            skip(t, p);
            if (t[p] == '}') {
                ++p;
                return j;
            }
            for (;;) {
                skip(t, p);
                const std::string key = parse_string(t, p);
                skip(t, p);
                if (t[p] != ':') throw std::invalid_argument("json shim: expected ':'");
                ++p;
                j.obj_[key] = parse_value(t, p);
                skip(t, p);
                if (t[p] == ',') {
                    ++p;
                    continue;
                }
                if (t[p] == '}') {
                    ++p;
                    break;
                }
                throw std::invalid_argument("json shim: expected ',' or '}'");
            }
        } else if (t[p] == '"') {
            j.k_ = kind::string, j.str_ = parse_string(t, p);
        } else if (t.compare(p, 4, "true") == 0) {
            j.k_ = kind::boolean, j.num_ = 1, p += 4;
        } else if (t.compare(p, 5, "false") == 0) {
            j.k_ = kind::boolean, j.num_ = 0, p += 5;
        } else if (t.compare(p, 4, "null") == 0) {
            p += 4;
        } else {
            char *end = nullptr;
            j.num_ = std::strtod(t.c_str() + p, &end);
            if (end == t.c_str() + p) throw std::invalid_argument("json shim: bad value");
            j.k_ = kind::number, p = static_cast<size_t>(end - t.c_str());
        }
        return j;
    }
    kind k_;
    std::string str_;
    double num_ = 0.0;
    std::map<std::string, json> obj_;
};

}// namespace nlohmann
