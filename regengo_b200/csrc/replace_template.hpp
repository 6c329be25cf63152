// Replace templates: `replace.Parse` (replace/template.go:60-163) and the lookups the generated
// ReplaceAllBytesAppend performs per segment (internal/compiler/replace.go:393-453).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rgx {

// One resolved segment: literal bytes, or the bytes of a capture group (0 = the whole match).  A reference to a
// group the pattern does not have (index out of range: CaptureByIndex returns nil, replace.go:50-52; unknown name:
// no case of the switch fires, replace.go:436-453) expands to nothing and is dropped here.
struct TemplateSegment {
  int group = -1;        // >= 0: capture group; -1: literal
  std::string literal;
};

// Go: unicode.IsLetter / unicode.IsDigit (category L / Nd), over the front-end's tables
bool rune_is_letter(int32_t r);
bool rune_is_digit(int32_t r);

// false + err ("at position %d: ...", the text of the reference's panic) on a malformed template
bool parse_replace_template(const std::string& tmpl, const std::vector<std::string>& capture_names,
                            std::vector<TemplateSegment>& out, std::string& err);

}  // namespace rgx
