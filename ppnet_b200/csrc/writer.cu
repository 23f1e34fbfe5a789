// N1: the dataset writer of MapGenerate.generate_map_randomly (EDaGe-PP/MapGenerate.py:144-149): one JSON line per map
// appended to ./unsolved_problems.txt, {"Index", "Init", "End", "Length", "Obstacles"}.  Host code (no kernel): formats the
// downloaded arrays straight into the file, byte for byte what Python's json.dumps writes -- floats as repr(float)
// (shortest round-trip digits; exponent form below 1e-4 and from 1e16; a trailing ".0" on integers), "NaN" / "Infinity".
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "common.cuh"

namespace ppnet {

// repr(float) of CPython (float_repr_style = 'short'): shortest digits that round-trip, formatted with decpt rules
static void py_float(std::string& out, double x) {
    if (std::isnan(x)) { out += "NaN"; return; }
    if (std::isinf(x)) { out += x > 0 ? "Infinity" : "-Infinity"; return; }
    if (x == 0.0) { out += std::signbit(x) ? "-0.0" : "0.0"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    *r.ptr = 0;
    const char* p = buf;
    if (*p == '-') { out += '-'; ++p; }
    const char* e = std::strchr(p, 'e');
    char digits[32];
    int nd = 0;
    for (const char* q = p; q < e; ++q) if (*q != '.') digits[nd++] = *q;
    const int exp10 = std::atoi(e + 1);
    const int decpt = exp10 + 1;                                  // position of the decimal point relative to the digits
    if (decpt > 16 || decpt < -3) {                               // exponent notation: d[.ddd]e[+-]XX (at least two exponent digits)
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
        char eb[16];
        std::snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out += eb;
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

}  // namespace ppnet

using namespace ppnet;

// problems [first, first + n): index[i] (the "Index" field), init / end [n][2], length[n], obs[n][omax][3] + obs_cnt[n].
// The lines are appended to `path` (append != 0) or replace it; `out_bytes` (may be NULL) receives the bytes written.
extern "C" int ppnet_write_problems_jsonl(const char* path, int32_t append, int64_t n, const int64_t* index, const double* init,
                                          const double* end, const double* length, const double* obs, const int32_t* obs_cnt,
                                          int32_t omax, int64_t* out_bytes) {
    PPNET_REQUIRE(path && n >= 0 && omax >= 0, "write_problems: bad arguments");
    PPNET_REQUIRE(n == 0 || (index && init && end && length && obs_cnt && (obs || omax == 0)), "write_problems: null pointer");
    std::string buf;
    buf.reserve((size_t)n * (128 + 60 * (size_t)omax));
    for (int64_t i = 0; i < n; ++i) {
        buf += "{\"Index\": ";
        buf += std::to_string((long long)index[i]);
        buf += ", \"Init\": [";
        py_float(buf, init[2 * i]); buf += ", "; py_float(buf, init[2 * i + 1]);
        buf += "], \"End\": [";
        py_float(buf, end[2 * i]); buf += ", "; py_float(buf, end[2 * i + 1]);
        buf += "], \"Length\": ";
        py_float(buf, length[i]);
        buf += ", \"Obstacles\": [";
        const int c = obs_cnt[i] < omax ? obs_cnt[i] : omax;
        for (int k = 0; k < c; ++k) {
            const double* o = obs + ((size_t)i * omax + k) * 3;
            buf += k ? ", [" : "[";
            py_float(buf, o[0]); buf += ", "; py_float(buf, o[1]); buf += ", "; py_float(buf, o[2]);
            buf += "]";
        }
        buf += "]}\n";
    }
    FILE* f = std::fopen(path, append ? "ab" : "wb");
    if (!f) { set_error("write_problems: cannot open %s", path); return PPNET_E_INVALID; }
    const size_t w = std::fwrite(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    if (w != buf.size()) { set_error("write_problems: short write to %s", path); return PPNET_E_INVALID; }
    if (out_bytes) *out_bytes = (int64_t)buf.size();
    return PPNET_OK;
}
