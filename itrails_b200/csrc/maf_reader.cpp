// maf_reader.cpp — MAF ingest for the hot path (host C++, no GPU, no Biopython).
//
// Replaces (reference paths relative to /root/reference/src/itrails):
//   maf_parser          read_data.py:94-117   -> symbols (uint16, 0..624) + block offsets
//   parse_coordinates   read_data.py:146-220  -> per-column reference coordinates / -9
//
// The reference goes through Biopython's AlignIO.parse(file, "maf") and resolves every
// column with an O(625) list.index.  Here the file is mmap'ed, block boundaries are
// found in one pass, blocks are parsed by a pool of threads, and a column is converted
// with a 256-entry byte table and base-5 arithmetic.  Semantics kept (the block iterator
// is Biopython 1.84's MafIterator, Bio/AlignIO/MafIO.py):
//   * outside a block a line whose first character is 'a' opens one (as many key=value
//     words as '=' signs, else an error); every other line (##maf, '#', track, blank) is
//     skipped;
//   * inside a block: first character 's' = a row "s src start size strand srcSize text"
//     with exactly 7 fields (else an error; strand '-' is -1, anything else +1; a '.' in the
//     text copies the letter of the block's FIRST row); 'i', 'e', 'q', '#' lines are
//     skipped; a blank line or the end of the file closes the block; any other line —
//     another 'a' line without a blank line before it, an indented row — is an error, and
//     so are rows of unequal length;  "\r\n" line ends are accepted;
//   * species = src up to the first '.'; rows of species outside the list are skipped;
//     if a species occurs twice the last row wins (dict assignment, read_data.py:108-109);
//   * symbols: block kept iff all four species are present (read_data.py:110);
//     '-' counts as 'N' (:109), letters are upper-cased (:114), any other character is an
//     error (list.index raises ValueError);
//   * coordinates: block kept iff exactly four rows belong to listed species (:171-173,
//     :181 — a duplicated species makes five and drops the block from the coordinate
//     list only); forward strand starts at `start`, reverse strand at srcSize - start and
//     runs backwards (:197-201, :213); gaps and blocks without the reference species give -9.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "../../include/itrails_b200.h"

namespace {

struct Row {
    const char *src;
    size_t src_len;
    int64_t start, src_size;
    int strand;
    const char *text;
    size_t text_len;
};

struct BlockOut {
    bool keep_sym = false, keep_coord = false;
    std::vector<uint16_t> sym;
    std::vector<int64_t> coord;
    std::deque<std::string> dots;     // rows whose '.' were replaced (stable addresses)
    std::string err;
};

struct Tables {
    uint8_t digit[256];       // A,C,T,G -> 0..3 (read_data.py:13 order), N and '-' -> 4, else 255
    uint16_t code_to_index[625];
    Tables() {
        memset(digit, 255, sizeof digit);
        const char *nuc = "ACTG";
        for (int i = 0; i < 4; ++i) {
            digit[(unsigned char)nuc[i]] = (uint8_t)i;
            digit[(unsigned char)(nuc[i] + 32)] = (uint8_t)i;
        }
        digit[(unsigned char)'N'] = digit[(unsigned char)'n'] = digit[(unsigned char)'-'] = 4;
        int next = 256;
        for (int a = 0; a < 5; ++a)
            for (int b = 0; b < 5; ++b)
                for (int c = 0; c < 5; ++c)
                    for (int d = 0; d < 5; ++d) {
                        const int code = ((a * 5 + b) * 5 + c) * 5 + d;
                        if (a < 4 && b < 4 && c < 4 && d < 4) code_to_index[code] = (uint16_t)(64 * a + 16 * b + 4 * c + d);
                        else code_to_index[code] = (uint16_t)next++;
                    }
    }
};
const Tables TAB;

inline const char *skip_ws(const char *p, const char *e) {
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    return p;
}
inline const char *skip_tok(const char *p, const char *e) {
    while (p < e && *p != ' ' && *p != '\t' && *p != '\r') ++p;
    return p;
}
bool parse_i64(const char *b, const char *e, int64_t *out) {       // Python int(): optional sign, digits
    bool neg = false;
    if (b < e && (*b == '+' || *b == '-')) neg = *b++ == '-';
    if (b == e) return false;
    int64_t v = 0;
    for (const char *p = b; p < e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    *out = neg ? -v : v;
    return true;
}

// Parses the "s" rows of one block [b, e): every line whose FIRST character is 's'
// (the boundary pass has already rejected lines that may not appear inside a block).
bool parse_rows(const char *b, const char *e, std::vector<Row> &rows, std::string &err) {
    rows.clear();
    const char *p = b;
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', e - p);
        const char *le = nl ? nl : e;
        if (p < le && *p == 's') {
            const char *f[8], *g[8];
            int n = 0;
            const char *t = p;
            while (n < 8) {
                t = skip_ws(t, le);
                if (t >= le || *t == '\r') break;
                f[n] = t;
                t = skip_tok(t, le);
                g[n] = t;
                ++n;
            }
            if (n != 7) {
                err = "Error parsing alignment - 's' line must have 7 fields: " + std::string(p, std::min<size_t>(60, le - p));
                return false;
            }
            Row r{};
            int64_t size_field = 0;
            r.src = f[1];
            r.src_len = g[1] - f[1];
            if (!parse_i64(f[2], g[2], &r.start) || !parse_i64(f[3], g[3], &size_field) || !parse_i64(f[5], g[5], &r.src_size)) {
                err = "invalid integer in MAF sequence line: " + std::string(p, std::min<size_t>(60, le - p));
                return false;
            }
            r.strand = (g[4] - f[4] == 1 && *f[4] == '-') ? -1 : 1;
            r.text = f[6];
            r.text_len = g[6] - f[6];
            rows.push_back(r);
        }
        p = nl ? nl + 1 : e;
    }
    return true;
}

inline int species_of(const Row &r, const char *const sp[4], const size_t sp_len[4]) {
    size_t n = 0;
    while (n < r.src_len && r.src[n] != '.') ++n;
    for (int k = 0; k < 4; ++k)
        if (n == sp_len[k] && memcmp(r.src, sp[k], n) == 0) return k;
    return -1;
}

void do_block(const char *b, const char *e, const char *const sp[4], const size_t sp_len[4], const char *ref,
              size_t ref_len, bool want_coord, BlockOut &out) {
    std::vector<Row> rows;
    if (!parse_rows(b, e, rows, out.err)) return;
    if (rows.empty()) return;
    // a '.' stands for the letter of the block's first row at that column (MafIterator);
    // the substituted text lives in out.dots for the lifetime of the block
    if (memchr(rows[0].text, '.', rows[0].text_len)) {
        out.err = "Found dot/period in first sequence of alignment";
        return;
    }
    for (size_t k = 1; k < rows.size(); ++k) {
        Row &r = rows[k];
        if (!memchr(r.text, '.', r.text_len)) continue;
        const size_t n = std::min(r.text_len, rows[0].text_len);      // (zip() stops at the shorter one)
        out.dots.emplace_back(r.text, n);
        std::string &t = out.dots.back();
        for (size_t i = 0; i < n; ++i)
            if (t[i] == '.') t[i] = rows[0].text[i];
        r.text = t.data();
        r.text_len = n;
    }
    const size_t len = rows[0].text_len;
    for (const Row &r : rows)
        if (r.text_len != len) {
            out.err = "Sequences must all be the same length";
            return;
        }
    const Row *pick[4] = {nullptr, nullptr, nullptr, nullptr};
    const Row *ref_row = nullptr;
    int acc = 0;
    for (const Row &r : rows) {
        const int k = species_of(r, sp, sp_len);
        if (k >= 0) {
            pick[k] = &r;
            ++acc;
        }
        if (want_coord) {
            size_t n = 0;
            while (n < r.src_len && r.src[n] != '.') ++n;
            if (n == ref_len && memcmp(r.src, ref, n) == 0) ref_row = &r;
        }
    }
    if (pick[0] && pick[1] && pick[2] && pick[3]) {
        out.keep_sym = true;
        out.sym.resize(len);
        for (size_t i = 0; i < len; ++i) {
            int code = 0;
            for (int k = 0; k < 4; ++k) {
                const uint8_t d = TAB.digit[(unsigned char)pick[k]->text[i]];
                if (d == 255) {
                    out.err = std::string("'") + pick[k]->text[i] + "' is not a valid nucleotide in a MAF column";
                    return;
                }
                code = code * 5 + d;
            }
            out.sym[i] = TAB.code_to_index[code];
        }
    }
    if (want_coord && acc == 4) {
        out.keep_coord = true;
        out.coord.assign(len, -9);
        if (ref_row) {
            int64_t st = ref_row->strand == 1 ? ref_row->start : ref_row->src_size - ref_row->start;
            for (size_t i = 0; i < len; ++i)
                if (ref_row->text[i] != '-') {
                    out.coord[i] = st;
                    st += ref_row->strand;
                }
        }
    }
}

}  // namespace

// The parsed blocks stay as the worker threads left them; `off` / `coord_off` are the offsets of
// the kept blocks in the concatenated outputs.  itr_maf_export copies them into the caller's
// arrays with all threads (the first touch of those pages is then parallel too: at 10 bytes per
// column the single-threaded concatenation + its page faults were half of the reader's time);
// the pointer accessors build the concatenation on first use.
struct itr_maf {
    std::vector<BlockOut> outs;
    std::vector<size_t> sym_blocks, coord_blocks;        // indices into outs of the kept blocks
    std::vector<uint16_t> sym;
    std::vector<int64_t> off{0};
    std::vector<int64_t> coord;
    std::vector<int64_t> coord_off{0};
    bool has_coord = false, flat_sym = false, flat_coord = false;
};

static void maf_copy_out(const itr_maf *m, uint16_t *sym_out, int64_t *coord_out, int n_threads) {
    const size_t ns = sym_out ? m->sym_blocks.size() : 0, nc = coord_out ? m->coord_blocks.size() : 0;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(nt, 1), ns + nc));
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= ns + nc) break;
            if (i < ns) {
                const BlockOut &o = m->outs[m->sym_blocks[i]];
                if (!o.sym.empty()) memcpy(sym_out + m->off[i], o.sym.data(), o.sym.size() * sizeof(uint16_t));
            } else {
                const BlockOut &o = m->outs[m->coord_blocks[i - ns]];
                if (!o.coord.empty()) memcpy(coord_out + m->coord_off[i - ns], o.coord.data(), o.coord.size() * sizeof(int64_t));
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

static int set_err(char *err, int cap, const std::string &msg, int code) {
    if (err && cap > 0) {
        strncpy(err, msg.c_str(), cap - 1);
        err[cap - 1] = 0;
    }
    return code;
}

extern "C" int itr_maf_read(const char *path, const char *const species[4], const char *ref, int n_threads,
                            itr_maf **out, char *err, int err_cap) {
    if (!path || !species || !out) return set_err(err, err_cap, "itr_maf_read: NULL argument", ITR_ERR_ARG);
    *out = nullptr;
    size_t sp_len[4];
    for (int k = 0; k < 4; ++k) {
        if (!species[k]) return set_err(err, err_cap, "itr_maf_read: four species names are required", ITR_ERR_ARG);
        sp_len[k] = strlen(species[k]);
    }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return set_err(err, err_cap, std::string("cannot open ") + path, ITR_ERR_ARG);
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return set_err(err, err_cap, std::string("cannot stat ") + path, ITR_ERR_ARG);
    }
    const size_t size = (size_t)st.st_size;
    itr_maf *m = new itr_maf();
    m->has_coord = ref != nullptr;
    if (size == 0) {
        close(fd);
        *out = m;
        return ITR_OK;
    }
    const char *base = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) {
        delete m;
        return set_err(err, err_cap, std::string("cannot mmap ") + path, ITR_ERR_NOMEM);
    }
    madvise((void *)base, size, MADV_SEQUENTIAL);
    const char *end = base + size;
    // block boundaries with MafIterator's state machine: [line after an "a" line, the blank
    // line that closes the block / end of file)
    std::vector<std::pair<const char *, const char *>> blocks;
    {
        const char *p = base, *cur = nullptr;
        std::string bad;
        while (p < end && bad.empty()) {
            const char *nl = (const char *)memchr(p, '\n', end - p);
            const char *le = nl ? nl : end;
            const char *q = skip_ws(p, le);
            const bool blank = (q == le) || (*q == '\r' && q + 1 == le);
            const char tag = p < le ? *p : '\0';
            if (cur) {
                if (blank) {
                    blocks.push_back({cur, p});
                    cur = nullptr;
                } else if (tag != 's' && tag != 'i' && tag != 'e' && tag != 'q' && tag != '#') {
                    bad = "Error parsing alignment - unexpected line: " + std::string(p, std::min<size_t>(60, le - p));
                }
            } else if (tag == 'a') {
                // as many whitespace-separated words after the first as '=' signs
                size_t words = 0, eqs = 0;
                bool in_word = false;
                for (const char *c = p; c < le; ++c) {
                    if (*c == '=') ++eqs;
                    const bool ws = (*c == ' ' || *c == '\t' || *c == '\r');
                    if (!ws && !in_word) ++words;
                    in_word = !ws;
                }
                if (words - 1 != eqs) bad = "Error parsing alignment - invalid key in 'a' line";
                cur = nl ? nl + 1 : end;
            }
            p = nl ? nl + 1 : end;
        }
        if (!bad.empty()) {
            munmap((void *)base, size);
            delete m;
            return set_err(err, err_cap, bad, ITR_ERR_ARG);
        }
        if (cur) blocks.push_back({cur, end});
    }
    const size_t nb = blocks.size();
    std::vector<BlockOut> &outs = m->outs;
    outs.resize(nb);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(nt, 1), (nb + 15) / 16));
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    const size_t ref_len = ref ? strlen(ref) : 0;
    auto worker = [&]() {
        for (;;) {
            const size_t i0 = next.fetch_add(16);
            if (i0 >= nb || failed.load()) break;
            for (size_t i = i0; i < std::min(nb, i0 + 16); ++i) {
                do_block(blocks[i].first, blocks[i].second, species, sp_len, ref, ref_len, ref != nullptr, outs[i]);
                if (!outs[i].err.empty()) failed.store(true);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    munmap((void *)base, size);
    for (size_t i = 0; i < nb; ++i)
        if (!outs[i].err.empty()) {       // the first failing block in file order, like a sequential parse
            const std::string msg = outs[i].err;
            delete m;
            return set_err(err, err_cap, msg, ITR_ERR_ARG);
        }
    int64_t n_sym = 0, n_coord = 0;
    for (size_t i = 0; i < nb; ++i) {
        BlockOut &o = outs[i];
        o.dots.clear();
        if (o.keep_sym) {
            m->sym_blocks.push_back(i);
            n_sym += (int64_t)o.sym.size();
            m->off.push_back(n_sym);
        }
        if (o.keep_coord) {
            m->coord_blocks.push_back(i);
            n_coord += (int64_t)o.coord.size();
            m->coord_off.push_back(n_coord);
        }
    }
    *out = m;
    return ITR_OK;
}

extern "C" void itr_maf_free(itr_maf *m) { delete m; }
extern "C" int64_t itr_maf_num_blocks(const itr_maf *m) { return m ? (int64_t)m->off.size() - 1 : 0; }
extern "C" int64_t itr_maf_num_columns(const itr_maf *m) { return m ? m->off.back() : 0; }
extern "C" int itr_maf_export(const itr_maf *m, uint16_t *sym_out, int64_t *coord_out, int n_threads) {
    if (!m) return ITR_ERR_ARG;
    if (coord_out && !m->has_coord) return ITR_ERR_STATE;
    maf_copy_out(m, sym_out, coord_out, n_threads);
    return ITR_OK;
}
extern "C" const uint16_t *itr_maf_symbols(const itr_maf *m) {
    if (!m) return nullptr;
    itr_maf *w = const_cast<itr_maf *>(m);
    if (!w->flat_sym) {
        w->sym.resize((size_t)m->off.back());
        maf_copy_out(m, w->sym.data(), nullptr, 0);
        w->flat_sym = true;
    }
    return m->sym.data();
}
extern "C" const int64_t *itr_maf_offsets(const itr_maf *m) { return m ? m->off.data() : nullptr; }
extern "C" int64_t itr_maf_num_coord_blocks(const itr_maf *m) { return m && m->has_coord ? (int64_t)m->coord_off.size() - 1 : 0; }
extern "C" const int64_t *itr_maf_coordinates(const itr_maf *m) {
    if (!m || !m->has_coord) return nullptr;
    itr_maf *w = const_cast<itr_maf *>(m);
    if (!w->flat_coord) {
        w->coord.resize((size_t)m->coord_off.back());
        maf_copy_out(m, nullptr, w->coord.data(), 0);
        w->flat_coord = true;
    }
    return m->coord.data();
}
extern "C" const int64_t *itr_maf_coord_offsets(const itr_maf *m) { return m && m->has_coord ? m->coord_off.data() : nullptr; }
