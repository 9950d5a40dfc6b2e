// miniproto_rt.hpp -- run-time support for the classes oracle/ref_shim/miniprotoc.py generates.
// TEST INFRASTRUCTURE (oracle/_ref): a stand-in for the slice of libprotobuf the reference's layer sources touch
// (RepeatedField/RepeatedPtrField accessors, float reflection, text format), because the image has no protobuf.
#ifndef MINIPROTO_RT_HPP_
#define MINIPROTO_RT_HPP_
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace google { namespace protobuf {

template <typename T>
class RepeatedField {
 public:
  int size() const { return (int)v_.size(); }
  T Get(int i) const { return v_.at(i); }
  void Set(int i, T x) { v_.at(i) = x; }
  void Add(T x) { v_.push_back(x); }
  void Clear() { v_.clear(); }
  const T* data() const { return v_.data(); }
  T* mutable_data() { return v_.data(); }
  void Reserve(int n) { v_.reserve(n); }
  void Resize(int n, T x) { v_.resize(n, x); }
  typename std::vector<T>::const_iterator begin() const { return v_.begin(); }
  typename std::vector<T>::const_iterator end() const { return v_.end(); }
  T operator[](int i) const { return v_[i]; }
 private:
  std::vector<T> v_;
};

template <typename T>
class RepeatedPtrField {
 public:
  RepeatedPtrField() {}
  RepeatedPtrField(const RepeatedPtrField& o) { *this = o; }
  RepeatedPtrField& operator=(const RepeatedPtrField& o) {
    if (this != &o) { v_.clear(); for (auto& p : o.v_) v_.emplace_back(new T(*p)); }
    return *this;
  }
  int size() const { return (int)v_.size(); }
  const T& Get(int i) const { return *v_.at(i); }
  T* Mutable(int i) { return v_.at(i).get(); }
  T* Add() { v_.emplace_back(new T()); return v_.back().get(); }
  void Clear() { v_.clear(); }
  const T& operator[](int i) const { return *v_[i]; }
 private:
  std::vector<std::unique_ptr<T>> v_;
};

inline void not_float() { throw std::runtime_error("miniproto: reflection only supports optional float fields"); }

struct FieldDescriptor {
  const char* name_;
  bool is_float_;
  float default_float_;
  int index_ = 0;
  FieldDescriptor(const char* n, bool f, float d) : name_(n), is_float_(f), default_float_(d) {}
  float default_value_float() const { if (!is_float_) not_float(); return default_float_; }
  const std::string name() const { return name_; }
  int index() const { return index_; }
};

class Descriptor {
 public:
  Descriptor(const char* name, FieldDescriptor* f, int n) : name_(name), f_(f), n_(n) {
    for (int i = 0; i < n; ++i) f[i].index_ = i;
  }
  int field_count() const { return n_; }
  const FieldDescriptor* field(int i) const { return &f_[i]; }
  const std::string name() const { return name_; }
 private:
  const char* name_; FieldDescriptor* f_; int n_;
};

class Message {
 public:
  virtual ~Message() {}
  virtual float _GetFloat(int idx) const = 0;
  virtual void _SetFloat(int idx, float v) = 0;
  virtual void _ClearField(int idx) = 0;
};

class Reflection {
 public:
  static const Reflection* get() { static Reflection r; return &r; }
  float GetFloat(const Message& m, const FieldDescriptor* f) const { return m._GetFloat(f->index_); }
  void SetFloat(Message* m, const FieldDescriptor* f, float v) const { m->_SetFloat(f->index_, v); }
  void ClearField(Message* m, const FieldDescriptor* f) const { m->_ClearField(f->index_); }
};

// ---- text format -------------------------------------------------------------------------------------------
class TextTok {
 public:
  explicit TextTok(const std::string& s) : s_(s), p_(0) { advance(); }
  [[noreturn]] void fail(const std::string& m) const {
    throw std::runtime_error("miniproto text format: " + m + " near offset " + std::to_string(p_));
  }
  // Reads the next field name; returns false at the closing token (consumed) or at end of input when close==nullptr.
  bool next_field(const char* close, std::string* name) {
    while (kind_ == ';' || kind_ == ',') advance();
    if (kind_ == 0) { if (close) fail("unexpected end of input"); return false; }
    if (close && kind_ == close[0]) { advance(); return false; }
    if (kind_ != 'i') fail("expected a field name, got '" + tok_ + "'");
    *name = tok_; advance(); return true;
  }
  void colon() { if (kind_ != ':') fail("expected ':'"); advance(); in_list_ = false; if (kind_ == '[') { in_list_ = true; advance(); } }
  bool list_more() {
    if (!in_list_) return false;
    if (kind_ == ',') { advance(); return true; }
    if (kind_ == ']') { advance(); in_list_ = false; return false; }
    fail("expected ',' or ']'");
  }
  const char* open_msg() {
    if (kind_ == ':') advance();
    if (kind_ == '{') { advance(); return "}"; }
    if (kind_ == '<') { advance(); return ">"; }
    fail("expected '{'");
  }
  std::string str() {
    if (kind_ != 's') fail("expected a string");
    std::string r;
    while (kind_ == 's') { r += tok_; advance(); }
    return r;
  }
  template <typename T> T num() {
    if (kind_ != 'n' && kind_ != 'i') fail("expected a number");
    std::string t = tok_; advance();
    if (t == "true" || t == "True" || t == "t") return (T)1;
    if (t == "false" || t == "False" || t == "f") return (T)0;
    if (!t.empty() && (t.back() == 'f' || t.back() == 'F') && t.find("inf") == std::string::npos) t.pop_back();
    char* e = nullptr;
    if (std::is_floating_point<T>::value) {
      // text format floats are parsed as double, then narrowed (protobuf's TextFormat does the same)
      double d = std::strtod(t.c_str(), &e);
      if (*e) fail("bad number '" + t + "'");
      return (T)d;
    }
    if (std::is_signed<T>::value) { long long v = std::strtoll(t.c_str(), &e, 0); if (*e) fail("bad integer '" + t + "'"); return (T)v; }
    unsigned long long v = std::strtoull(t.c_str(), &e, 0); if (*e) fail("bad integer '" + t + "'"); return (T)v;
  }
  template <typename E> void enumv(bool (*parse)(const std::string&, E*), E* out) {
    if (kind_ == 'n') { *out = (E)std::atoi(tok_.c_str()); advance(); return; }
    if (kind_ != 'i' || !parse(tok_, out)) fail("bad enum value '" + tok_ + "'");
    advance();
  }
 private:
  void advance() {
    const std::string& s = s_;
    for (;;) {
      while (p_ < s.size() && (s[p_] == ' ' || s[p_] == '\t' || s[p_] == '\n' || s[p_] == '\r')) ++p_;
      if (p_ < s.size() && s[p_] == '#') { while (p_ < s.size() && s[p_] != '\n') ++p_; continue; }
      break;
    }
    tok_.clear();
    if (p_ >= s.size()) { kind_ = 0; return; }
    char c = s[p_];
    if (c == '"' || c == '\'') {
      kind_ = 's'; ++p_;
      while (p_ < s.size() && s[p_] != c) {
        if (s[p_] == '\\' && p_ + 1 < s.size()) {
          char e = s[p_ + 1]; p_ += 2;
          switch (e) { case 'n': tok_ += '\n'; break; case 't': tok_ += '\t'; break; case 'r': tok_ += '\r'; break;
                       case '0': tok_ += '\0'; break; default: tok_ += e; }
        } else tok_ += s[p_++];
      }
      if (p_ >= s.size()) fail("unterminated string");
      ++p_; return;
    }
    if (std::isalpha((unsigned char)c) || c == '_') {
      kind_ = 'i';
      while (p_ < s.size() && (std::isalnum((unsigned char)s[p_]) || s[p_] == '_' || s[p_] == '.')) tok_ += s[p_++];
      return;
    }
    if (std::isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
      kind_ = 'n';
      tok_ += s[p_++];
      while (p_ < s.size() && (std::isalnum((unsigned char)s[p_]) || s[p_] == '.' ||
             ((s[p_] == '-' || s[p_] == '+') && (s[p_ - 1] == 'e' || s[p_ - 1] == 'E')))) tok_ += s[p_++];
      return;
    }
    kind_ = c; tok_ = std::string(1, c); ++p_;
  }
  const std::string s_;
  size_t p_;
  int kind_;          // 0 end, 'i' identifier, 'n' number, 's' string, else the punctuation character
  std::string tok_;
  bool in_list_ = false;
};

template <typename M>
inline bool ParseTextInto(const std::string& text, M* m) { TextTok t(text); m->ParseText(t, nullptr); return true; }

}}  // namespace google::protobuf
#endif
