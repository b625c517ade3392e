// Minimal stand-in for <boost/algorithm/string.hpp>: only what the reference's
// LoadGraph / SAM parsing uses (split + is_any_of + token_compress_on).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build); Boost is not installed in this image.
#pragma once
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
namespace boost {
struct any_of_pred { std::string chars; bool operator()(char c) const { return chars.find(c) != std::string::npos; } };
inline any_of_pred is_any_of(const char* s) { return any_of_pred{std::string(s)}; }
enum token_compress_mode_type { token_compress_on, token_compress_off };
template <class Seq, class Pred>
Seq& split(Seq& out, const std::string& in, Pred pred, token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool last_was_sep = false;
  for (char c : in) {
    if (pred(c)) {
      if (mode == token_compress_on && last_was_sep) continue;
      out.push_back(cur); cur.clear(); last_was_sep = true;
    } else { cur.push_back(c); last_was_sep = false; }
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost
