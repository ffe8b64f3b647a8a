// Minimal JSON DOM reader for `.crtscene` files (schema: SURVEY.md App. D; reference loader SceneParser.cpp:39-66,
// which uses rapidjson -- a third-party dependency the reference does not vendor).  Numeric arrays -- >99.9 % of a
// scene file -- are stored as flat double vectors.  Numbers are converted like rapidjson's default (non
// full-precision) path for the inputs our scenes contain: <= 15 significant digits with a small exponent are
// exactly representable products/quotients (correctly rounded, same as strtod); anything else falls back to strtod.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace crt::json {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, NumArray, Object } kind = Null;
  bool b = false;
  double num = 0;
  bool isInt = false;  // written without '.', 'e' (rapidjson IsInt/IsUint vs IsFloat distinction)
  std::string str;
  std::vector<Value> arr;
  std::vector<double> nums;  // NumArray
  std::vector<std::pair<std::string, Value>> obj;

  const Value *find(const char *key) const {
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool isArray() const { return kind == Array || kind == NumArray; }
  size_t size() const { return kind == NumArray ? nums.size() : arr.size(); }
  double number(size_t i) const { return kind == NumArray ? nums[i] : arr[i].num; }
};

class Parser {
 public:
  Parser(const char *b, const char *e) : p(b), end(e) {}
  Value parse() {
    Value v = value();
    ws();
    return v;
  }

 private:
  const char *p, *end;
  [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("crtscene JSON: ") + m); }
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
  }
  static double pow10i(int e) {
    static const double t[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                               1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[e];
  }
  double number(bool &isInt) {
    const char *s = p;
    bool neg = false;
    if (p < end && *p == '-') {
      neg = true;
      p++;
    }
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool simple = true;
    isInt = true;
    while (p < end && *p >= '0' && *p <= '9') {
      if (digits < 18) {
        mant = mant * 10 + (*p - '0');
        if (mant) digits++;
      } else
        simple = false;
      p++;
    }
    if (p < end && *p == '.') {
      isInt = false;
      p++;
      while (p < end && *p >= '0' && *p <= '9') {
        if (digits < 18) {
          mant = mant * 10 + (*p - '0');
          if (mant) digits++;
          exp10--;
        } else
          simple = false;
        p++;
      }
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
      isInt = false;
      p++;
      bool eneg = false;
      if (p < end && (*p == '+' || *p == '-')) {
        eneg = *p == '-';
        p++;
      }
      int ev = 0;
      while (p < end && *p >= '0' && *p <= '9') {
        if (ev < 10000) ev = ev * 10 + (*p - '0');
        p++;
      }
      exp10 += eneg ? -ev : ev;
    }
    if (p == s) fail("bad number");
    if (simple && digits <= 15 && exp10 >= -22 && exp10 <= 22) {
      double d = static_cast<double>(mant);
      if (exp10 < 0)
        d /= pow10i(-exp10);
      else if (exp10 > 0)
        d *= pow10i(exp10);
      return neg ? -d : d;
    }
    std::string tmp(s, p);
    return std::strtod(tmp.c_str(), nullptr);
  }
  std::string string() {
    std::string out;
    p++;  // opening quote
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        p++;
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = 0;
            for (int i = 0; i < 4 && p + 1 < end; i++) {
              p++;
              cp = cp * 16 + (*p <= '9' ? *p - '0' : ((*p | 32) - 'a' + 10));
            }
            if (cp < 0x80)
              out += static_cast<char>(cp);
            else if (cp < 0x800) {
              out += static_cast<char>(0xC0 | (cp >> 6));
              out += static_cast<char>(0x80 | (cp & 63));
            } else {
              out += static_cast<char>(0xE0 | (cp >> 12));
              out += static_cast<char>(0x80 | ((cp >> 6) & 63));
              out += static_cast<char>(0x80 | (cp & 63));
            }
            break;
          }
          default: out += *p;
        }
        p++;
      } else
        out += *p++;
    }
    if (p >= end) fail("unterminated string");
    p++;
    return out;
  }
  Value value() {
    ws();
    if (p >= end) fail("unexpected end");
    Value v;
    char c = *p;
    if (c == '{') {
      v.kind = Value::Object;
      p++;
      ws();
      if (p < end && *p == '}') {
        p++;
        return v;
      }
      while (true) {
        ws();
        if (p >= end || *p != '"') fail("expected key");
        std::string k = string();
        ws();
        if (p >= end || *p != ':') fail("expected ':'");
        p++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p < end && *p == ',') {
          p++;
          continue;
        }
        if (p < end && *p == '}') {
          p++;
          break;
        }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      p++;
      ws();
      if (p < end && *p == ']') {
        p++;
        v.kind = Value::NumArray;
        return v;
      }
      ws();
      if (p < end && (*p == '-' || (*p >= '0' && *p <= '9'))) {
        // flat numeric array fast path
        v.kind = Value::NumArray;
        while (true) {
          ws();
          bool isInt;
          if (!(p < end && (*p == '-' || (*p >= '0' && *p <= '9')))) fail("mixed array");
          v.nums.push_back(number(isInt));
          ws();
          if (p < end && *p == ',') {
            p++;
            continue;
          }
          if (p < end && *p == ']') {
            p++;
            break;
          }
          fail("expected ',' or ']'");
        }
      } else {
        v.kind = Value::Array;
        while (true) {
          v.arr.push_back(value());
          ws();
          if (p < end && *p == ',') {
            p++;
            continue;
          }
          if (p < end && *p == ']') {
            p++;
            break;
          }
          fail("expected ',' or ']'");
        }
      }
    } else if (c == '"') {
      v.kind = Value::String;
      v.str = string();
    } else if (c == 't' && end - p >= 4 && !std::memcmp(p, "true", 4)) {
      v.kind = Value::Bool;
      v.b = true;
      p += 4;
    } else if (c == 'f' && end - p >= 5 && !std::memcmp(p, "false", 5)) {
      v.kind = Value::Bool;
      v.b = false;
      p += 5;
    } else if (c == 'n' && end - p >= 4 && !std::memcmp(p, "null", 4)) {
      v.kind = Value::Null;
      p += 4;
    } else {
      v.kind = Value::Number;
      v.num = number(v.isInt);
    }
    return v;
  }
};

}  // namespace crt::json
