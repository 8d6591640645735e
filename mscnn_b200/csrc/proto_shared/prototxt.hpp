// Protobuf text-format reader (schema-less) for Caffe .prototxt files.
//
// The reference parses net definitions with google::protobuf::TextFormat
// (src/caffe/util/io.cpp:34-44 ReadProtoFromTextFile).  libprotobuf is not available, so this
// is an independent reader of the same grammar into a generic tree; typed views
// (caffe::LayerParameter etc.) are filled from the tree by proto.hpp.  Supported: nested
// messages with `name { }` and `name: { }`, scalars (numbers incl. negative / exponent, bare
// enum identifiers, booleans, quoted strings with escapes and adjacent-string concatenation),
// repeated fields by repetition or `[a, b]` lists, several fields per line, optional `,` / `;`
// separators and `#` comments anywhere (all of which occur in the shipped
// examples/*/mscnn_deploy.prototxt files).
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace prototxt {

struct Node;
struct Field {
  std::string name;
  bool is_message = false;
  std::string scalar;           // raw token (strings unescaped, without quotes)
  bool quoted = false;
  std::shared_ptr<Node> message;
  int line = 0;
};

struct Node {
  std::vector<Field> fields;  // in file order

  std::vector<const Field*> all(const std::string& name) const {
    std::vector<const Field*> r;
    for (const Field& f : fields)
      if (f.name == name) r.push_back(&f);
    return r;
  }
  const Field* first(const std::string& name) const {
    for (const Field& f : fields)
      if (f.name == name) return &f;
    return nullptr;
  }
  bool has(const std::string& name) const { return first(name) != nullptr; }
  const Node* child(const std::string& name) const {
    const Field* f = first(name);
    return (f && f->is_message) ? f->message.get() : nullptr;
  }
  std::string str(const std::string& name, const std::string& dflt = "") const {
    const Field* f = first(name);
    return (f && !f->is_message) ? f->scalar : dflt;
  }
  double num(const std::string& name, double dflt) const {
    const Field* f = first(name);
    return (f && !f->is_message) ? to_double(*f) : dflt;
  }
  bool boolean(const std::string& name, bool dflt) const {
    const Field* f = first(name);
    if (!f || f->is_message) return dflt;
    if (f->scalar == "true" || f->scalar == "True" || f->scalar == "1" || f->scalar == "t") return true;
    if (f->scalar == "false" || f->scalar == "False" || f->scalar == "0" || f->scalar == "f") return false;
    throw std::runtime_error("prototxt: line " + std::to_string(f->line) + ": bad bool '" + f->scalar + "'");
  }
  std::vector<double> nums(const std::string& name) const {
    std::vector<double> r;
    for (const Field* f : all(name))
      if (!f->is_message) r.push_back(to_double(*f));
    return r;
  }
  std::vector<std::string> strs(const std::string& name) const {
    std::vector<std::string> r;
    for (const Field* f : all(name))
      if (!f->is_message) r.push_back(f->scalar);
    return r;
  }
  static double to_double(const Field& f) {
    // text format allows a trailing 'f' on floats and inf/nan spellings
    std::string s = f.scalar;
    if (s.size() > 1 && (s.back() == 'f' || s.back() == 'F') && s.find_first_of("xX") == std::string::npos)
      s.pop_back();
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0')
      throw std::runtime_error("prototxt: line " + std::to_string(f.line) + ": bad number '" + f.scalar +
                               "' for field " + f.name);
    return v;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& text) : s_(text) {}

  std::shared_ptr<Node> parse() {
    auto root = std::make_shared<Node>();
    parse_fields(*root, /*closing=*/'\0');
    return root;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  int line_ = 1;

  [[noreturn]] void fail(const std::string& what) const {
    throw std::runtime_error("prototxt: line " + std::to_string(line_) + ": " + what);
  }
  void skip_ws() {
    while (i_ < s_.size()) {
      const char c = s_[i_];
      if (c == '\n') { ++line_; ++i_; }
      else if (std::isspace(static_cast<unsigned char>(c)) || c == ',' || c == ';') ++i_;
      else if (c == '#') { while (i_ < s_.size() && s_[i_] != '\n') ++i_; }
      else break;
    }
  }
  static bool ident_char(char c) { return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '.'; }

  std::string parse_ident() {
    const size_t b = i_;
    while (i_ < s_.size() && ident_char(s_[i_])) ++i_;
    if (b == i_) fail(std::string("expected a field name, found '") + (i_ < s_.size() ? s_[i_] : '$') + "'");
    return s_.substr(b, i_ - b);
  }

  std::string parse_quoted() {
    std::string out;
    for (;;) {  // adjacent string literals concatenate
      const char q = s_[i_++];
      while (true) {
        if (i_ >= s_.size()) fail("unterminated string");
        char c = s_[i_++];
        if (c == q) break;
        if (c == '\n') ++line_;
        if (c == '\\' && i_ < s_.size()) {
          const char e = s_[i_++];
          switch (e) {
            case 'n': c = '\n'; break;
            case 't': c = '\t'; break;
            case 'r': c = '\r'; break;
            case '\\': c = '\\'; break;
            case '\'': c = '\''; break;
            case '"': c = '"'; break;
            default: out.push_back('\\'); c = e; break;
          }
        }
        out.push_back(c);
      }
      skip_ws();
      if (i_ < s_.size() && (s_[i_] == '"' || s_[i_] == '\'')) continue;
      break;
    }
    return out;
  }

  void parse_scalar_into(Node& node, const std::string& name) {
    Field f;
    f.name = name;
    f.line = line_;
    if (s_[i_] == '"' || s_[i_] == '\'') {
      f.quoted = true;
      f.scalar = parse_quoted();
    } else {
      const size_t b = i_;
      while (i_ < s_.size() && (ident_char(s_[i_]) || s_[i_] == '-' || s_[i_] == '+')) ++i_;
      if (b == i_) fail("expected a value for field '" + name + "'");
      f.scalar = s_.substr(b, i_ - b);
    }
    node.fields.push_back(std::move(f));
  }

  void parse_fields(Node& node, char closing) {
    for (;;) {
      skip_ws();
      if (i_ >= s_.size()) {
        if (closing != '\0') fail("unexpected end of input, missing '}'");
        return;
      }
      if (s_[i_] == '}' || s_[i_] == '>') {
        if (closing == '\0') fail("unbalanced '}'");
        ++i_;
        return;
      }
      const std::string name = parse_ident();
      skip_ws();
      bool colon = false;
      if (i_ < s_.size() && s_[i_] == ':') { colon = true; ++i_; skip_ws(); }
      if (i_ >= s_.size()) fail("unexpected end of input after '" + name + "'");
      if (s_[i_] == '{' || s_[i_] == '<') {
        ++i_;
        Field f;
        f.name = name;
        f.is_message = true;
        f.line = line_;
        f.message = std::make_shared<Node>();
        parse_fields(*f.message, '}');
        node.fields.push_back(std::move(f));
      } else if (s_[i_] == '[') {
        if (!colon) fail("list value needs ':'");
        ++i_;
        for (;;) {
          skip_ws();
          if (i_ >= s_.size()) fail("unterminated list");
          if (s_[i_] == ']') { ++i_; break; }
          parse_scalar_into(node, name);
        }
      } else {
        if (!colon) fail("expected ':' or '{' after '" + name + "'");
        parse_scalar_into(node, name);
      }
    }
  }
};

inline std::shared_ptr<Node> parse_string(const std::string& text) { return Parser(text).parse(); }

inline std::shared_ptr<Node> parse_file(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("prototxt: cannot open " + path);
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();
  return parse_string(text);
}

}  // namespace prototxt
