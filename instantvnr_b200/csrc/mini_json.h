// mini_json.h -- a small JSON value + parser (// and /* */ comments allowed, as the
// reference parses its model files with nlohmann's ignore_comments=true, api.cpp:20) and a
// BSON reader/writer for the reference's params.json blobs (core/network.cu:827-939).
// Header-only, host-only, no dependencies.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mj {

struct Value;
using Object = std::vector<std::pair<std::string, Value>>;   // keeps insertion order
using Array = std::vector<Value>;

struct Value {
  enum Type { Null, Bool, Int, Double, String, ArrayT, ObjectT, Binary } type = Null;
  bool b = false;
  int64_t i = 0;
  double d = 0;
  std::string s;                 // String, or raw bytes for Binary
  uint8_t subtype = 0;           // Binary subtype
  std::shared_ptr<Array> arr;
  std::shared_ptr<Object> obj;

  Value() {}
  static Value make_object() { Value v; v.type = ObjectT; v.obj = std::make_shared<Object>(); return v; }
  static Value make_array() { Value v; v.type = ArrayT; v.arr = std::make_shared<Array>(); return v; }
  static Value from(double x) { Value v; v.type = Double; v.d = x; return v; }
  static Value from_int(int64_t x) { Value v; v.type = Int; v.i = x; return v; }
  static Value from(bool x) { Value v; v.type = Bool; v.b = x; return v; }
  static Value from(const std::string& x) { Value v; v.type = String; v.s = x; return v; }
  static Value binary(const void* p, size_t n) { Value v; v.type = Binary; v.s.assign((const char*)p, n); return v; }

  bool is_object() const { return type == ObjectT; }
  bool is_number() const { return type == Int || type == Double; }
  bool is_string() const { return type == String; }
  bool contains(const std::string& k) const {
    if (type != ObjectT) return false;
    for (auto& kv : *obj) if (kv.first == k) return true;
    return false;
  }
  const Value& at(const std::string& k) const {
    if (type != ObjectT) throw std::runtime_error("json: not an object (key '" + k + "')");
    for (auto& kv : *obj) if (kv.first == k) return kv.second;
    throw std::runtime_error("json: missing key '" + k + "'");
  }
  Value& set(const std::string& k, const Value& v) {
    if (type != ObjectT) { *this = make_object(); }
    for (auto& kv : *obj) if (kv.first == k) { kv.second = v; return kv.second; }
    obj->emplace_back(k, v);
    return obj->back().second;
  }
  double num() const {
    if (type == Int) return (double)i;
    if (type == Double) return d;
    if (type == Bool) return b ? 1 : 0;
    throw std::runtime_error("json: not a number");
  }
  double value(const std::string& k, double def) const { return contains(k) ? at(k).num() : def; }
  std::string value(const std::string& k, const char* def) const { return contains(k) && at(k).is_string() ? at(k).s : std::string(def); }
  Value value_obj(const std::string& k) const { return contains(k) ? at(k) : make_object(); }
};

class Parser {
  const char* p; const char* e;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json parse error: ") + m); }
  void ws() {
    for (;;) {
      while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
      if (p + 1 < e && p[0] == '/' && p[1] == '/') { while (p < e && *p != '\n') ++p; continue; }
      if (p + 1 < e && p[0] == '/' && p[1] == '*') { p += 2; while (p + 1 < e && !(p[0] == '*' && p[1] == '/')) ++p; p += 2; continue; }
      break;
    }
  }
  std::string str() {
    if (*p != '"') fail("expected string");
    ++p; std::string out;
    while (p < e && *p != '"') {
      if (*p == '\\') {
        ++p; if (p >= e) fail("bad escape");
        switch (*p) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break;
          case 'u': { if (p + 4 >= e) fail("bad \\u"); unsigned c = (unsigned)std::stoul(std::string(p + 1, p + 5), nullptr, 16); p += 4;
                      if (c < 0x80) out += (char)c; else if (c < 0x800) { out += (char)(0xC0 | (c >> 6)); out += (char)(0x80 | (c & 0x3F)); }
                      else { out += (char)(0xE0 | (c >> 12)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); } break; }
          default: out += *p;
        }
        ++p;
      } else out += *p++;
    }
    if (p >= e) fail("unterminated string");
    ++p; return out;
  }
  Value val() {
    ws(); if (p >= e) fail("unexpected end");
    if (*p == '{') {
      ++p; Value v = Value::make_object(); ws();
      if (p < e && *p == '}') { ++p; return v; }
      for (;;) {
        ws(); std::string k = str(); ws(); if (p >= e || *p != ':') fail("expected ':'"); ++p;
        v.obj->emplace_back(k, val()); ws();
        if (p < e && *p == ',') { ++p; continue; }
        if (p < e && *p == '}') { ++p; return v; }
        fail("expected ',' or '}'");
      }
    }
    if (*p == '[') {
      ++p; Value v = Value::make_array(); ws();
      if (p < e && *p == ']') { ++p; return v; }
      for (;;) {
        v.arr->push_back(val()); ws();
        if (p < e && *p == ',') { ++p; continue; }
        if (p < e && *p == ']') { ++p; return v; }
        fail("expected ',' or ']'");
      }
    }
    if (*p == '"') return Value::from(str());
    if (!strncmp(p, "true", 4) && e - p >= 4) { p += 4; return Value::from(true); }
    if (!strncmp(p, "false", 5) && e - p >= 5) { p += 5; return Value::from(false); }
    if (!strncmp(p, "null", 4) && e - p >= 4) { p += 4; return Value(); }
    const char* s = p; bool isd = false;
    if (p < e && (*p == '-' || *p == '+')) ++p;
    while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) { if (*p == '.' || *p == 'e' || *p == 'E') isd = true; ++p; }
    if (s == p) fail("unexpected character");
    std::string t(s, p);
    return isd ? Value::from(std::stod(t)) : Value::from_int(std::stoll(t));
  }
 public:
  static Value parse(const std::string& text) {
    Parser ps; ps.p = text.data(); ps.e = text.data() + text.size();
    Value v = ps.val(); ps.ws();
    if (ps.p != ps.e) ps.fail("trailing characters");
    return v;
  }
};

inline void dump(const Value& v, std::string& out) {
  switch (v.type) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Int: out += std::to_string(v.i); break;
    case Value::Double: { char buf[40]; snprintf(buf, sizeof buf, "%.17g", v.d); out += buf; break; }
    case Value::String: out += '"'; for (char c : v.s) { if (c == '"' || c == '\\') out += '\\'; out += c; } out += '"'; break;
    case Value::Binary: out += "\"<binary>\""; break;
    case Value::ArrayT: { out += '['; bool f = true; for (auto& x : *v.arr) { if (!f) out += ','; f = false; dump(x, out); } out += ']'; break; }
    case Value::ObjectT: { out += '{'; bool f = true; for (auto& kv : *v.obj) { if (!f) out += ','; f = false; out += '"' + kv.first + "\":"; dump(kv.second, out); } out += '}'; break; }
  }
}

// ------------------------------- BSON ---------------------------------------
// Subset produced by nlohmann::json::to_bson for the reference's params files:
// double (0x01), string (0x02), document (0x03), array (0x04), binary (0x05),
// bool (0x08), null (0x0A), int32 (0x10), int64 (0x12).
namespace bson {
inline void put32(std::string& o, int32_t v) { o.append((const char*)&v, 4); }
inline void write_doc(const Value& v, std::string& o, bool as_array);
inline void write_elem(const std::string& key, const Value& v, std::string& o) {
  auto head = [&](uint8_t t) { o += (char)t; o += key; o += '\0'; };
  switch (v.type) {
    case Value::Null: head(0x0A); break;
    case Value::Bool: head(0x08); o += (char)(v.b ? 1 : 0); break;
    case Value::Int:
      if (v.i >= INT32_MIN && v.i <= INT32_MAX) { head(0x10); put32(o, (int32_t)v.i); }
      else { head(0x12); o.append((const char*)&v.i, 8); }
      break;
    case Value::Double: head(0x01); o.append((const char*)&v.d, 8); break;
    case Value::String: head(0x02); put32(o, (int32_t)v.s.size() + 1); o += v.s; o += '\0'; break;
    case Value::Binary: head(0x05); put32(o, (int32_t)v.s.size()); o += (char)v.subtype; o += v.s; break;
    case Value::ObjectT: head(0x03); write_doc(v, o, false); break;
    case Value::ArrayT: head(0x04); write_doc(v, o, true); break;
  }
}
inline void write_doc(const Value& v, std::string& o, bool as_array) {
  size_t start = o.size(); put32(o, 0);
  if (as_array) { size_t k = 0; for (auto& x : *v.arr) write_elem(std::to_string(k++), x, o); }
  else for (auto& kv : *v.obj) write_elem(kv.first, kv.second, o);
  o += '\0';
  int32_t len = (int32_t)(o.size() - start); memcpy(&o[start], &len, 4);
}
inline std::string write(const Value& root) { std::string o; write_doc(root, o, false); return o; }

struct Reader {
  const uint8_t* p; const uint8_t* e;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("bson parse error: ") + m); }
  int32_t i32() { if (e - p < 4) fail("truncated"); int32_t v; memcpy(&v, p, 4); p += 4; return v; }
  std::string cstr() { const uint8_t* s = p; while (p < e && *p) ++p; if (p >= e) fail("truncated key"); std::string k((const char*)s, p - s); ++p; return k; }
  Value doc(bool as_array) {
    const uint8_t* start = p; int32_t len = i32();
    if (len < 5 || start + len > e) fail("bad document length");
    const uint8_t* end = start + len;
    Value v = as_array ? Value::make_array() : Value::make_object();
    while (p < end - 1) {
      uint8_t t = *p++; std::string k = cstr(); Value x;
      switch (t) {
        case 0x01: { if (e - p < 8) fail("truncated"); double d; memcpy(&d, p, 8); p += 8; x = Value::from(d); break; }
        case 0x02: { int32_t n = i32(); if (n < 1 || e - p < n) fail("bad string"); x = Value::from(std::string((const char*)p, n - 1)); p += n; break; }
        case 0x03: x = doc(false); break;
        case 0x04: x = doc(true); break;
        case 0x05: { int32_t n = i32(); if (n < 0 || e - p < n + 1) fail("bad binary"); uint8_t st = *p++; x = Value::binary(p, n); x.subtype = st; p += n; break; }
        case 0x08: { if (p >= e) fail("truncated"); x = Value::from(*p++ != 0); break; }
        case 0x0A: break;
        case 0x10: x = Value::from_int(i32()); break;
        case 0x12: { if (e - p < 8) fail("truncated"); int64_t q; memcpy(&q, p, 8); p += 8; x = Value::from_int(q); break; }
        default: fail("unsupported element type");
      }
      if (as_array) v.arr->push_back(x); else v.obj->emplace_back(k, x);
    }
    if (p != end - 1 || *p != 0) fail("missing terminator");
    ++p; return v;
  }
};
inline Value read(const void* data, size_t n) { Reader r; r.p = (const uint8_t*)data; r.e = r.p + n; return r.doc(false); }
}  // namespace bson

}  // namespace mj
