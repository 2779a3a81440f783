// Host-side ingest of the reference's xyz dialect (no CUDA): replaces the Python parsing loops of
// charge_gn.gen_padded_init_state (reference charge_gn.py:301-338) with a multi-threaded C++ reader that
// fills the packed SoA arrays epnn_infer_batch takes (offsets, xyz, species, Q).
//
// Dialect (charge_gn.py:309-330):
//   line 1: atom count -- ignored (atoms = every non-blank line after line 2)
//   line 2: first token = net charge Q (float32); the rest is ignored
//   line 3+: "Elem x y z [ignored columns]"; numbers are parsed as float64 and rounded once to float32, exactly like
//            numpy's np.array(list_of_strings, dtype=np.float32)
// Element symbols index the table chosen by n_x (9: H C N O F S Cl Br, infer.py:13-30; 10: H C N O F P S Cl Br,
// charge_gn.py:9-28); an unknown symbol is an error (the reference raises KeyError, charge_gn.py:326-327).
#include <locale.h>
#include <cmath>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/epnn_b200.h"

namespace {

struct ParsedFile {
    float Q = 0.f;
    std::vector<float> xyz;
    std::vector<int32_t> species;
    int error = 0;              // 0 ok, 1 cannot open, 2 malformed, 3 unknown element
    std::string bad;            // offending token / message
};

const char* const TABLE10[] = {"H", "C", "N", "O", "F", "P", "S", "Cl", "Br"};
const char* const TABLE9[] = {"H", "C", "N", "O", "F", "S", "Cl", "Br"};

int species_of(const char* sym, size_t len, int n_x) {
    const char* const* tab = n_x == 9 ? TABLE9 : TABLE10;
    const int n = n_x - 1;
    for (int i = 0; i < n; ++i)
        if (strlen(tab[i]) == len && memcmp(tab[i], sym, len) == 0) return i;
    return -1;
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// next whitespace-delimited token of [p, end); returns false if the line holds no more tokens
bool next_token(const char*& p, const char* end, const char*& tok, size_t& len) {
    while (p < end && is_space(*p)) ++p;
    if (p >= end) return false;
    tok = p;
    while (p < end && !is_space(*p)) ++p;
    len = (size_t)(p - tok);
    return true;
}

bool parse_float32(const char* tok, size_t len, float& out) {
    char buf[64];
    if (len == 0 || len >= sizeof buf) return false;
    memcpy(buf, tok, len);
    buf[len] = 0;
    for (size_t i = 0; i < len; ++i)        // hex floats ("0x1p3") are not part of the dialect (Python's float() rejects them too)
        if (buf[i] == 'x' || buf[i] == 'X' || buf[i] == 'p' || buf[i] == 'P') return false;
    static const locale_t c_loc = newlocale(LC_NUMERIC_MASK, "C", (locale_t)0);      // '.' is the decimal point whatever the process locale says
    char* e = nullptr;
    const double d = c_loc ? strtod_l(buf, &e, c_loc) : strtod(buf, &e);       // correctly rounded float64, then one rounding to float32 (numpy semantics)
    if (e != buf + len || !std::isfinite(d)) return false;                     // a nan / inf coordinate or charge is an input error, not a number
    out = (float)d;
    return true;
}

void parse_text(const char* text, size_t size, int n_x, ParsedFile& out) {
    const char* p = text;
    const char* const end = text + size;
    int line_no = 0;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = eol ? eol : end;
        ++line_no;
        const char* q = p;
        const char* tok; size_t len;
        if (line_no == 2) {
            if (!next_token(q, le, tok, len) || !parse_float32(tok, len, out.Q)) { out.error = 2; out.bad = "line 2: net charge"; return; }
        } else if (line_no > 2) {
            if (next_token(q, le, tok, len)) {          // blank lines are skipped
                const int sp = species_of(tok, len, n_x);
                if (sp < 0) { out.error = 3; out.bad.assign(tok, len); return; }
                float c[3];
                for (int k = 0; k < 3; ++k) {
                    const char* t2; size_t l2;
                    if (!next_token(q, le, t2, l2) || !parse_float32(t2, l2, c[k])) { out.error = 2; out.bad = "atom line " + std::to_string(line_no); return; }
                }
                out.species.push_back(sp);
                out.xyz.push_back(c[0]); out.xyz.push_back(c[1]); out.xyz.push_back(c[2]);
            }
        }
        p = eol ? eol + 1 : end;
    }
    if (line_no < 3 || out.species.empty()) { out.error = 2; out.bad = "no atom lines"; }
}

void parse_path(const char* path, int n_x, ParsedFile& out) {
    FILE* f = fopen(path, "rb");
    if (!f) { out.error = 1; out.bad = path; return; }
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    parse_text(text.data(), text.size(), n_x, out);
}

}  // namespace

struct epnn_xyz_batch {
    std::vector<int32_t> offsets;
    std::vector<float> xyz;
    std::vector<int32_t> species;
    std::vector<float> Q;
    int error = 0, error_file = -1;
    std::string message;
};

extern "C" int epnn_xyz_load(const char* const* paths, int64_t n_files, int n_x, int threads, epnn_xyz_batch** out) {
    if (!out) return EPNN_E_INVALID;
    *out = nullptr;
    if (n_files < 0 || (n_files > 0 && !paths) || (n_x != 9 && n_x != 10)) return EPNN_E_INVALID;
    epnn_xyz_batch* b = new (std::nothrow) epnn_xyz_batch();
    if (!b) return EPNN_E_NOMEM;
    std::vector<ParsedFile> files((size_t)n_files);
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > n_files) nt = (int)(n_files > 0 ? n_files : 1);
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_files) break;
            parse_path(paths[i], n_x, files[(size_t)i]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    b->offsets.assign((size_t)n_files + 1, 0);
    int64_t total = 0;                                   // the packed offsets are int32 (the C-ABI of epnn_infer_batch): refuse to wrap
    for (int64_t i = 0; i < n_files; ++i) total += (int64_t)files[(size_t)i].species.size();
    if (total > 0x7fffffffLL) {
        b->error = 2; b->error_file = -1;
        b->message = "more than 2^31 - 1 atoms in one batch (" + std::to_string(total) + "); load the directory in several batches";
        *out = b;
        return EPNN_E_UNSUPPORTED;
    }
    for (int64_t i = 0; i < n_files; ++i) {
        const ParsedFile& pf = files[(size_t)i];
        if (pf.error && !b->error) {
            b->error = pf.error; b->error_file = (int)i;
            b->message = (pf.error == 1 ? "cannot open " : pf.error == 3 ? "unknown element '" : "malformed xyz (") + pf.bad +
                         (pf.error == 3 ? "'" : pf.error == 2 ? ")" : "") + (pf.error == 1 ? "" : std::string(" in ") + paths[i]);
        }
        b->offsets[(size_t)i + 1] = b->offsets[(size_t)i] + (int32_t)pf.species.size();
    }
    if (!b->error) {
        b->xyz.reserve((size_t)b->offsets.back() * 3);
        b->species.reserve((size_t)b->offsets.back());
        b->Q.reserve((size_t)n_files);
        for (const ParsedFile& pf : files) {
            b->xyz.insert(b->xyz.end(), pf.xyz.begin(), pf.xyz.end());
            b->species.insert(b->species.end(), pf.species.begin(), pf.species.end());
            b->Q.push_back(pf.Q);
        }
    }
    *out = b;
    return b->error == 0 ? EPNN_OK : EPNN_E_INVALID;
}

extern "C" int epnn_xyz_parse_text(const char* text, size_t len, int n_x, epnn_xyz_batch** out) {
    if (!out || (!text && len) || (n_x != 9 && n_x != 10)) return EPNN_E_INVALID;
    epnn_xyz_batch* b = new (std::nothrow) epnn_xyz_batch();
    if (!b) return EPNN_E_NOMEM;
    ParsedFile pf;
    parse_text(text, len, n_x, pf);
    b->offsets = {0, (int32_t)pf.species.size()};
    if (pf.error) {
        b->error = pf.error; b->error_file = 0;
        b->message = pf.error == 3 ? "unknown element '" + pf.bad + "'" : "malformed xyz (" + pf.bad + ")";
    } else {
        b->xyz = pf.xyz; b->species = pf.species; b->Q = {pf.Q};
    }
    *out = b;
    return b->error == 0 ? EPNN_OK : EPNN_E_INVALID;
}

extern "C" int64_t epnn_xyz_n_systems(const epnn_xyz_batch* b) { return b ? (int64_t)b->offsets.size() - 1 : 0; }
extern "C" int64_t epnn_xyz_n_atoms(const epnn_xyz_batch* b) { return b && !b->offsets.empty() ? b->offsets.back() : 0; }
extern "C" const int32_t* epnn_xyz_offsets(const epnn_xyz_batch* b) { return b ? b->offsets.data() : nullptr; }
extern "C" const float* epnn_xyz_coords(const epnn_xyz_batch* b) { return b ? b->xyz.data() : nullptr; }
extern "C" const int32_t* epnn_xyz_species(const epnn_xyz_batch* b) { return b ? b->species.data() : nullptr; }
extern "C" const float* epnn_xyz_charges(const epnn_xyz_batch* b) { return b ? b->Q.data() : nullptr; }
extern "C" int epnn_xyz_error(const epnn_xyz_batch* b, int* kind, int* file_index) {
    if (!b) return EPNN_E_INVALID;
    if (kind) *kind = b->error;
    if (file_index) *file_index = b->error_file;
    return EPNN_OK;
}
extern "C" const char* epnn_xyz_error_message(const epnn_xyz_batch* b) { return b ? b->message.c_str() : ""; }
extern "C" void epnn_xyz_free(epnn_xyz_batch* b) { delete b; }
