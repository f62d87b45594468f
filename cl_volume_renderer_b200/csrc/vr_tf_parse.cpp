// vr_tf_parse.cpp — strict parser / formatter for the run-time generated `is_event_gen` source.
//
// The reference JIT-compiles this text into every kernel that needs the transfer function
// (app/renderer.cpp:39,42; opencl_wrapper/include/clw_function.hpp:74-86).  Here the text is parsed into a clause
// table that the CUDA kernels evaluate directly, so a TF edit costs a table upload instead of a recompile.
//
// Accepted forms (anything else is VR_ERR_PARSE — there is no silent fallback):
//   (A) app/ui.cpp:160-168 + app/tf_part.cpp:55-79
//         inline bool is_event_gen(short value, short gradient, int4 *color){
//           if(value >= <min_v> && value <= <max_v>[ && gradient > <min_g> && gradient < <max_g>])
//          {
//             int4 tmp_color = {<r>,<g>,<b>,<a>};
//             *color = tmp_color;
//             return true;
//          }
//           ... more clauses ...
//           return false;
//         }
//   (B) tests/sdf/sdf_test.cpp:22, app/sdf_benchmark.cpp:18
//         inline bool is_event_gen(short value, short gradient, uint4 *color){ return (value > <K>); }
#include <cctype>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include "vr_internal.h"

namespace {

struct Scanner {
  const char* s;
  size_t i = 0;
  explicit Scanner(const char* src) : s(src) {}
  void ws() {
    while (s[i] && isspace((unsigned char)s[i])) ++i;
  }
  // match a literal token sequence, ignoring whitespace differences between tokens
  bool lit(const char* t) {
    size_t save = i;
    for (const char* p = t; *p;) {
      if (isspace((unsigned char)*p)) { ++p; ws(); continue; }
      ws();
      if (s[i] != *p) { i = save; return false; }
      ++i; ++p;
    }
    return true;
  }
  bool number(double* out) {
    ws();
    char* end = nullptr;
    double v = strtod(s + i, &end);
    if (end == s + i) return false;
    i = (size_t)(end - s);
    *out = v;
    return true;
  }
  bool integer(int* out) {
    ws();
    char* end = nullptr;
    long v = strtol(s + i, &end, 10);
    if (end == s + i) return false;
    i = (size_t)(end - s);
    *out = (int)v;
    return true;
  }
};

}  // namespace

extern "C" int vr_tf_parse(const char* src, vr_tf_rect* out, int cap, int* n) {
  VR_REQUIRE(src && out && n && cap > 0, "vr_tf_parse: null argument");
  Scanner sc(src);
  *n = 0;
  if (!sc.lit("inline bool is_event_gen ( short value , short gradient ,")) {
    vr_set_error("vr_tf_parse: missing `inline bool is_event_gen(short value, short gradient, ...` header");
    return VR_ERR_PARSE;
  }
  if (!sc.lit("int4 * color ) {") && !sc.lit("uint4 * color ) {")) {
    vr_set_error("vr_tf_parse: third parameter must be `int4 *color` or `uint4 *color`");
    return VR_ERR_PARSE;
  }
  for (;;) {
    double a, b, c, d;
    if (sc.lit("return false ; }")) break;
    if (sc.lit("return ( value >")) {  // form (B)
      if (!sc.number(&a) || !sc.lit(") ; }")) {
        vr_set_error("vr_tf_parse: malformed `return (value > K);`");
        return VR_ERR_PARSE;
      }
      if (*n >= cap) { vr_set_error("vr_tf_parse: more than %d clauses", cap); return VR_ERR_PARSE; }
      vr_tf_rect q{};
      q.min_v = (float)a;
      q.flags = VR_TF_THRESHOLD;
      out[(*n)++] = q;
      break;
    }
    if (!sc.lit("if ( value >=") || !sc.number(&a) || !sc.lit("&& value <=") || !sc.number(&b)) {
      vr_set_error("vr_tf_parse: expected `if(value >= A && value <= B` near offset %zu", sc.i);
      return VR_ERR_PARSE;
    }
    vr_tf_rect q{};
    q.min_v = (float)a;
    q.max_v = (float)b;
    if (sc.lit("&& gradient >")) {
      if (!sc.number(&c) || !sc.lit("&& gradient <") || !sc.number(&d)) {
        vr_set_error("vr_tf_parse: malformed gradient clause near offset %zu", sc.i);
        return VR_ERR_PARSE;
      }
      q.min_g = (float)c;
      q.max_g = (float)d;
      q.flags |= VR_TF_USE_GRADIENT;
    }
    int col[4];
    if (!sc.lit(") { int4 tmp_color = {") || !sc.integer(&col[0]) || !sc.lit(",") || !sc.integer(&col[1]) ||
        !sc.lit(",") || !sc.integer(&col[2]) || !sc.lit(",") || !sc.integer(&col[3]) ||
        !sc.lit("} ; * color = tmp_color ; return true ; }")) {
      vr_set_error("vr_tf_parse: malformed clause body near offset %zu", sc.i);
      return VR_ERR_PARSE;
    }
    for (int k = 0; k < 4; ++k) q.rgba[k] = col[k];
    if (*n >= cap) { vr_set_error("vr_tf_parse: more than %d clauses", cap); return VR_ERR_PARSE; }
    out[(*n)++] = q;
  }
  sc.ws();
  if (sc.s[sc.i] != '\0') {
    vr_set_error("vr_tf_parse: trailing text after is_event_gen at offset %zu", sc.i);
    return VR_ERR_PARSE;
  }
  return VR_OK;
}

// Text exactly as app/ui.cpp:160-168 / app/tf_part.cpp:55-79 emit it (default ostream float formatting).
extern "C" int vr_tf_format(const vr_tf_rect* rects, int n, char* out, size_t cap) {
  VR_REQUIRE(out && cap > 0 && (rects || n == 0), "vr_tf_format: null argument");
  std::ostringstream cl;
  if (n == 1 && (rects[0].flags & VR_TF_THRESHOLD)) {
    cl << "inline bool is_event_gen(short value, short gradient, uint4 *color){ return (value > " << rects[0].min_v
       << "); }";
  } else {
    cl << "inline bool is_event_gen(short value, short gradient, int4 *color){\n";
    for (int i = 0; i < n; ++i) {
      const vr_tf_rect& q = rects[i];
      VR_REQUIRE(!(q.flags & VR_TF_THRESHOLD), "vr_tf_format: threshold clause must be the only clause");
      cl << "  if(value >= " << q.min_v << " && value <= " << q.max_v;
      if (q.flags & VR_TF_USE_GRADIENT) cl << " && gradient > " << q.min_g << " && gradient < " << q.max_g;
      cl << ")\n {\n";
      cl << "    int4 tmp_color = {" << q.rgba[0] << "," << q.rgba[1] << "," << q.rgba[2] << "," << q.rgba[3] << "};\n";
      cl << "    *color = tmp_color;\n    return true;\n }\n";
    }
    cl << "  \n  return false;\n}\n";
  }
  const std::string s = cl.str();
  VR_REQUIRE(s.size() + 1 <= cap, "vr_tf_format: buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  return VR_OK;
}

TfTable vr_make_tf_table(const vr_tf_rect* rects, int n) {
  TfTable t{};
  t.n = n;
  t.needs_gradient = 0;
  for (int i = 0; i < n; ++i) {
    t.r[i] = rects[i];
    if (rects[i].flags & VR_TF_USE_GRADIENT) t.needs_gradient = 1;
    // (float)colour / 255.0f as ray_marching.cl:43-45,53,69-71 evaluate it per sample: one IEEE division, done once here
    for (int k = 0; k < 4; ++k) {
      volatile float c = (float)rects[i].rgba[k];
      volatile float q = c / 255.0f;
      t.e[i][k] = q;
    }
  }
  return t;
}
