// See csv_writer.h.  Host C++ only (no CUDA): the text side of itrails-posterior.
#include "csv_writer.h"

#include <atomic>
#include <charconv>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "../../include/itrails_b200.h"

namespace itr {

// Python's float repr (Objects/floatobject.c float_repr -> PyOS_double_to_string(x, 'r', 0,
// Py_DTSF_ADD_DOT_0)): shortest digit string that round-trips; with decpt = position of the
// decimal point relative to the digits, exponent notation iff decpt <= -4 or decpt > 16.
// std::to_chars(scientific) yields the same shortest digits (closest of the shortest).
int format_repr(double x, char *out) {
    char *p = out;
    if (std::isnan(x)) { std::memcpy(p, "nan", 3); return 3; }
    if (std::signbit(x)) { *p++ = '-'; x = -x; }
    if (std::isinf(x)) { std::memcpy(p, "inf", 3); return int(p - out) + 3; }
    if (x == 0.0) { std::memcpy(p, "0.0", 3); return int(p - out) + 3; }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);
    // sci = d[.ddd]e[+-]XX[X]
    char digits[24];
    int nd = 0;
    const char *q = sci;
    for (; q < r.ptr && *q != 'e'; ++q)
        if (*q != '.') digits[nd++] = *q;
    ++q;                                   // 'e'
    const bool neg = (*q == '-');
    ++q;
    int e = 0;
    for (; q < r.ptr; ++q) e = e * 10 + (*q - '0');
    if (neg) e = -e;
    const int decpt = e + 1;
    if (decpt <= -4 || decpt > 16) {
        *p++ = digits[0];
        if (nd > 1) {
            *p++ = '.';
            std::memcpy(p, digits + 1, nd - 1);
            p += nd - 1;
        }
        *p++ = 'e';
        int ee = decpt - 1;
        *p++ = ee < 0 ? '-' : '+';
        if (ee < 0) ee = -ee;
        if (ee >= 100) { *p++ = char('0' + ee / 100); ee %= 100; }
        *p++ = char('0' + ee / 10);
        *p++ = char('0' + ee % 10);
    } else if (decpt <= 0) {
        *p++ = '0';
        *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        std::memcpy(p, digits, nd);
        p += nd;
    } else if (decpt >= nd) {
        std::memcpy(p, digits, nd);
        p += nd;
        for (int i = nd; i < decpt; ++i) *p++ = '0';
        *p++ = '.';
        *p++ = '0';
    } else {
        std::memcpy(p, digits, decpt);
        p += decpt;
        *p++ = '.';
        std::memcpy(p, digits + decpt, nd - decpt);
        p += nd - decpt;
    }
    return int(p - out);
}

static inline char *put_i64(char *p, int64_t v) {
    auto r = std::to_chars(p, p + 24, v);
    return r.ptr;
}

bool PosteriorCsv::open(const char *path, int K, int n_threads, std::string &err, bool header) {
    close();
    fh_ = std::fopen(path, "wb");
    if (!fh_) { err = std::string("cannot open ") + path + " for writing"; return false; }
    std::setvbuf(fh_, nullptr, _IOFBF, 1 << 22);
    K_ = K;
    if (n_threads <= 0) n_threads = int(std::thread::hardware_concurrency());
    n_threads_ = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    std::string h = "alignment_block_idx,position_idx";
    for (int i = 0; i < K; ++i) h += ",prob_state_" + std::to_string(i);
    h += "\r\n";
    if (!header) h.clear();
    bytes_ = (int64_t)h.size();
    if (std::fwrite(h.data(), 1, h.size(), fh_) != h.size()) { err = "write failed"; return false; }
    return true;
}

bool PosteriorCsv::write_block(int64_t block_idx, const int64_t *positions, const double *post, int64_t n_rows,
                               std::string &err) {
    if (!fh_) { err = "posterior CSV is not open"; return false; }
    if (n_rows <= 0) return true;
    const int64_t ROWS = 4096;
    const int64_t n_chunks = (n_rows + ROWS - 1) / ROWS;
    if ((int64_t)chunks_.size() < n_chunks) chunks_.resize(n_chunks);
    const int K = K_;
    const size_t row_cap = 48 + (size_t)K * 26;
    std::atomic<int64_t> next{0};
    std::vector<char> ready(n_chunks, 0);
    std::mutex mu;
    std::condition_variable cv;
    auto worker = [&]() {
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= n_chunks) return;
            const int64_t r0 = c * ROWS, r1 = std::min(n_rows, r0 + ROWS);
            std::vector<char> &buf = chunks_[c];
            buf.resize((size_t)(r1 - r0) * row_cap);
            char *p = buf.data();
            for (int64_t r = r0; r < r1; ++r) {
                p = put_i64(p, block_idx);
                *p++ = ',';
                p = put_i64(p, positions ? positions[r] : r);
                const double *row = post + (size_t)r * K;
                for (int k = 0; k < K; ++k) {
                    *p++ = ',';
                    p += format_repr(row[k], p);
                }
                *p++ = '\r';
                *p++ = '\n';
            }
            buf.resize((size_t)(p - buf.data()));
            {
                std::lock_guard<std::mutex> lk(mu);
                ready[c] = 1;
            }
            cv.notify_all();
        }
    };
    const int nt = (int)std::min<int64_t>(n_threads_, n_chunks);
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    bool ok = true;
    for (int64_t c = 0; c < n_chunks; ++c) {       // the caller's thread writes in order
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready[c] != 0; });
        }
        if (ok && std::fwrite(chunks_[c].data(), 1, chunks_[c].size(), fh_) != chunks_[c].size()) ok = false;
        bytes_ += (int64_t)chunks_[c].size();
    }
    for (auto &t : pool) t.join();
    if (!ok) err = "write failed (disk full?)";
    return ok;
}

bool PosteriorCsv::close() {
    bool ok = true;
    if (fh_) { ok = std::fclose(fh_) == 0; fh_ = nullptr; }
    chunks_.clear();
    chunks_.shrink_to_fit();
    return ok;
}

}  // namespace itr

// ---- host-only C entry points (usable without a GPU context) -----------------------
extern "C" int itr_csv_format_double(double x, char *out, int cap) {
    if (!out || cap < 32) return -1;
    const int n = itr::format_repr(x, out);
    out[n] = 0;
    return n;
}

extern "C" int itr_csv_posterior_host(const char *path, int K, int64_t n_blocks, const int64_t *offsets,
                                      const int64_t *positions, const double *post, int n_threads) {
    return itr_csv_posterior_host_ex(path, K, n_blocks, offsets, positions, post, nullptr, 1, nullptr, n_threads);
}

extern "C" int itr_csv_posterior_host_ex(const char *path, int K, int64_t n_blocks, const int64_t *offsets,
                                         const int64_t *positions, const double *post, const int64_t *block_ids,
                                         int write_header, int64_t *block_bytes, int n_threads) {
    if (!path || K <= 0 || n_blocks < 0 || (n_blocks > 0 && (!offsets || !post))) return ITR_ERR_ARG;
    itr::PosteriorCsv w;
    std::string err;
    if (!w.open(path, K, n_threads, err, write_header != 0)) return ITR_ERR_IO;
    for (int64_t i = 0; i < n_blocks; ++i) {
        const int64_t c0 = offsets[i], n = offsets[i + 1] - c0;
        if (n < 0) return ITR_ERR_ARG;
        const int64_t before = w.bytes_written();
        if (!w.write_block(block_ids ? block_ids[i] : i, positions ? positions + c0 : nullptr, post + (size_t)c0 * K, n, err))
            return ITR_ERR_IO;
        if (block_bytes) block_bytes[i] = w.bytes_written() - before;
    }
    return w.close() ? ITR_OK : ITR_ERR_IO;
}
