#include "replace_template.hpp"

namespace rgx {

namespace {

// template.go:241-248 -- isNameStart / isNameContinue
bool name_start(int32_t r) { return r == '_' || rune_is_letter(r); }
bool name_continue(int32_t r) { return r == '_' || rune_is_letter(r) || rune_is_digit(r); }

// `for i, r := range s`: UTF-8 decoding with Go's rules (an invalid sequence yields U+FFFD and advances one byte)
int32_t decode_rune(const std::string& s, size_t i, size_t* width) {
  const auto b = [&](size_t k) -> uint32_t { return (uint8_t)s[k]; };
  const uint32_t c = b(i);
  *width = 1;
  if (c < 0x80) return (int32_t)c;
  size_t need;
  uint32_t r, lo;
  if (c >= 0xC2 && c <= 0xDF) { need = 1; r = c & 0x1F; lo = 0x80; }
  else if (c >= 0xE0 && c <= 0xEF) { need = 2; r = c & 0x0F; lo = 0x800; }
  else if (c >= 0xF0 && c <= 0xF4) { need = 3; r = c & 0x07; lo = 0x10000; }
  else return 0xFFFD;
  for (size_t k = 1; k <= need; k++) {
    if (i + k >= s.size() || (b(i + k) & 0xC0) != 0x80) return 0xFFFD;
    r = (r << 6) | (b(i + k) & 0x3F);
  }
  if (r < lo || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) return 0xFFFD;
  *width = need + 1;
  return (int32_t)r;
}

// template.go:251-267
bool valid_identifier(const std::string& s) {
  if (s.empty()) return false;
  size_t i = 0;
  bool first = true;
  while (i < s.size()) {
    size_t w;
    const int32_t r = decode_rune(s, i, &w);
    if (first ? !name_start(r) : !name_continue(r)) return false;
    first = false;
    i += w;
  }
  return true;
}

void add_group(std::vector<TemplateSegment>& out, size_t index, size_t n_groups) {
  if (index >= n_groups) return;   // CaptureByIndex: default -> nil
  TemplateSegment g;
  g.group = (int)index;
  out.push_back(g);
}

void add_named(std::vector<TemplateSegment>& out, const std::string& name, const std::vector<std::string>& names) {
  // replace.go:439-447: one case per NAMED group i >= 1; an unknown name appends nothing
  for (size_t i = 1; i < names.size(); i++)
    if (!names[i].empty() && names[i] == name) { add_group(out, i, names.size()); return; }
}

void add_literal(std::vector<TemplateSegment>& out, const std::string& lit) {
  if (lit.empty()) return;
  if (!out.empty() && out.back().group < 0) { out.back().literal += lit; return; }   // (adjacent literals append back to back anyway)
  TemplateSegment l;
  l.literal = lit;
  out.push_back(l);
}

}  // namespace

bool parse_replace_template(const std::string& t, const std::vector<std::string>& names, std::vector<TemplateSegment>& out,
                            std::string& err) {
  out.clear();
  const size_t n = t.size(), n_groups = names.size();
  size_t i = 0, literal_start = 0;
  while (i < n) {
    if (t[i] != '$') { i++; continue; }
    if (i > literal_start) add_literal(out, t.substr(literal_start, i - literal_start));
    if (i + 1 >= n) {               // "$" at the end: a literal dollar
      add_literal(out, "$");
      i++; literal_start = i;
      continue;
    }
    const uint8_t next = (uint8_t)t[i + 1];
    if (next == '$') {
      add_literal(out, "$");
      i += 2; literal_start = i;
    } else if (next == '{') {
      // parseBracedRef (template.go:166-208) on t[i:]
      const size_t close = t.find('}', i);
      const std::string at = "at position " + std::to_string(i) + ": ";
      if (close == std::string::npos) { err = at + "unclosed ${"; return false; }
      const std::string content = t.substr(i + 2, close - (i + 2));
      if (content.empty()) { err = at + "empty ${}"; return false; }
      if (content[0] >= '0' && content[0] <= '9') {
        size_t index = 0;
        for (char ch : content) {
          if (ch < '0' || ch > '9') { err = at + "invalid capture reference ${" + content + "}: mixed digits and non-digits"; return false; }
          index = index > (1u << 24) ? index : index * 10 + (size_t)(ch - '0');   // (saturates: any such index is out of range)
        }
        add_group(out, index, n_groups);   // ${0}: the whole match
      } else {
        if (!valid_identifier(content)) { err = at + "invalid capture name ${" + content + "}"; return false; }
        add_named(out, content, names);
      }
      i = close + 1; literal_start = i;
    } else if (next == '0') {
      add_group(out, 0, n_groups);
      i += 2; literal_start = i;
    } else if (next >= '1' && next <= '9') {
      // parseIndexedRef: one digit, or two when another digit follows
      size_t index = next - '0', consumed = 2;
      if (i + 2 < n && t[i + 2] >= '0' && t[i + 2] <= '9') { index = index * 10 + (size_t)(t[i + 2] - '0'); consumed = 3; }
      add_group(out, index, n_groups);
      i += consumed; literal_start = i;
    } else if (name_start((int32_t)next)) {
      // parseNamedRef: rune(s[end]) of a BYTE -- a byte >= 0x80 is read as the Latin-1 code point of that value
      size_t end = i + 2;
      while (end < n && name_continue((int32_t)(uint8_t)t[end])) end++;
      add_named(out, t.substr(i + 1, end - (i + 1)), names);
      i = end; literal_start = i;
    } else {
      add_literal(out, "$");
      i++; literal_start = i;
    }
  }
  if (i > literal_start) add_literal(out, t.substr(literal_start, i - literal_start));
  return true;
}

}  // namespace rgx
