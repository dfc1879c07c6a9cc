// wf_json.hpp — a small JSON reader for the WeldFormFEM input decks (examples/input/*.json of the reference).
// The reference parses its decks with the third-party nlohmann/json (vendored under include/common/nlohmann); this
// is an independent recursive-descent reader with the access pattern main.C uses: `j["Key"]` on a missing key gives
// a null value, readValue / readVector / readArray leave their target untouched on null (include/common/Input.h:32-98),
// `value(key, default)` returns the default when the key is absent (main.C:92-102).
#pragma once

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace wf_b200 {

class Json {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };
  Type type = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

  bool is_null() const { return type == Null; }
  bool contains(const std::string &k) const {
    for (auto &kv : obj)
      if (kv.first == k) return true;
    return false;
  }
  const Json &operator[](const std::string &k) const {
    static const Json nul;
    for (auto &kv : obj)
      if (kv.first == k) return kv.second;
    return nul;
  }
  const Json &operator[](const char *k) const { return (*this)[std::string(k)]; }
  const Json &operator[](size_t i) const {
    static const Json nul;
    return (type == Array && i < arr.size()) ? arr[i] : nul;
  }
  const Json &operator[](int i) const { return (*this)[(size_t)i]; }
  size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }
  double value(const std::string &k, double dflt) const {
    const Json &q = (*this)[k];
    return q.type == Number ? q.num : dflt;
  }

  static Json parse(const std::string &text) {
    Parser p{text, 0};
    Json j = p.value();
    p.ws();
    if (p.i != text.size()) p.fail("trailing characters");
    return j;
  }
  static Json parse_file(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
  }

 private:
  struct Parser {
    const std::string &s;
    size_t i;
    [[noreturn]] void fail(const std::string &what) const {
      size_t line = 1;
      for (size_t q = 0; q < i && q < s.size(); q++)
        if (s[q] == '\n') line++;
      throw std::runtime_error("JSON: " + what + " at line " + std::to_string(line));
    }
    void ws() {
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    Json value() {
      ws();
      if (i >= s.size()) fail("unexpected end");
      char c = s[i];
      Json j;
      if (c == '{') {
        j.type = Object;
        i++;
        ws();
        if (i < s.size() && s[i] == '}') { i++; return j; }
        for (;;) {
          ws();
          if (i >= s.size() || s[i] != '"') fail("expected a key");
          std::string k = string();
          ws();
          if (i >= s.size() || s[i] != ':') fail("expected ':'");
          i++;
          j.obj.emplace_back(k, value());
          ws();
          if (i < s.size() && s[i] == ',') { i++; continue; }
          if (i < s.size() && s[i] == '}') { i++; break; }
          fail("expected ',' or '}'");
        }
      } else if (c == '[') {
        j.type = Array;
        i++;
        ws();
        if (i < s.size() && s[i] == ']') { i++; return j; }
        for (;;) {
          j.arr.push_back(value());
          ws();
          if (i < s.size() && s[i] == ',') { i++; continue; }
          if (i < s.size() && s[i] == ']') { i++; break; }
          fail("expected ',' or ']'");
        }
      } else if (c == '"') {
        j.type = String;
        j.str = string();
      } else if (s.compare(i, 4, "true") == 0) { j.type = Bool; j.b = true; i += 4; }
      else if (s.compare(i, 5, "false") == 0) { j.type = Bool; j.b = false; i += 5; }
      else if (s.compare(i, 4, "null") == 0) { i += 4; }
      else {
        const char *b = s.c_str() + i;
        char *e = nullptr;
        double v = strtod(b, &e);
        if (e == b) fail(std::string("unexpected character '") + c + "'");
        j.type = Number;
        j.num = v;
        i += (size_t)(e - b);
      }
      return j;
    }
    std::string string() {
      std::string out;
      i++;  // opening quote
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\' && i + 1 < s.size()) {
          char e = s[i + 1];
          switch (e) {
            case 'n': out += '\n'; break;
            case 't': out += '\t'; break;
            case 'r': out += '\r'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'u': {  // basic-plane escape, kept as UTF-8
              unsigned cp = (unsigned)strtoul(s.substr(i + 2, 4).c_str(), nullptr, 16);
              if (cp < 0x80) out += (char)cp;
              else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
              else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
              i += 4;
              break;
            }
            default: out += e;
          }
          i += 2;
        } else out += s[i++];
      }
      if (i >= s.size()) fail("unterminated string");
      i++;
      return out;
    }
  };
};

// include/common/Input.h:32-98: targets stay untouched when the value is null
inline bool readValue(const Json &j, double &v) { if (j.is_null()) return false; if (j.type != Json::Number) throw std::runtime_error("JSON: number expected"); v = j.num; return true; }
inline bool readValue(const Json &j, int &v) { if (j.is_null()) return false; if (j.type != Json::Number) throw std::runtime_error("JSON: number expected"); v = (int)j.num; return true; }
inline bool readValue(const Json &j, bool &v) { if (j.is_null()) return false; if (j.type != Json::Bool) throw std::runtime_error("JSON: bool expected"); v = j.b; return true; }
inline bool readValue(const Json &j, std::string &v) { if (j.is_null()) return false; if (j.type != Json::String) throw std::runtime_error("JSON: string expected"); v = j.str; return true; }
inline bool readArray(const Json &j, std::vector<double> &v) {
  if (j.is_null()) return false;
  if (j.type != Json::Array) throw std::runtime_error("JSON: array expected");
  v.resize(j.arr.size());
  for (size_t i = 0; i < j.arr.size(); i++) v[i] = j.arr[i].num;
  return true;
}

}  // namespace wf_b200
