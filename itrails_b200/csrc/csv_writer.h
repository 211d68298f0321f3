// Host-side writer of `{prefix}.posterior.csv`, byte-identical to what the reference's
// `csv.writer` loop produces (workflow_posterior.py:697-716): "\r\n" line ends, ints as
// decimal, floats as Python's repr() (shortest round-trip digits, fixed notation for
// decimal exponents -4..15, else d.ddde±XX).  Rows are formatted by a pool of threads.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace itr {

// Python repr() of a double into out (>= 32 bytes); returns the length (no terminator).
int format_repr(double x, char *out);

class PosteriorCsv {
public:
    ~PosteriorCsv() { close(); }
    // Creates/truncates the file and writes the header line for K states (header = false:
    // a part file of a sharded run, spliced behind another file's header later).
    bool open(const char *path, int K, int n_threads, std::string &err, bool header = true);
    // Appends the rows of one block.  positions == nullptr writes 0..n_rows-1.
    bool write_block(int64_t block_idx, const int64_t *positions, const double *post, int64_t n_rows,
                     std::string &err);
    bool close();
    int64_t bytes_written() const { return bytes_; }

private:
    FILE *fh_ = nullptr;
    int K_ = 0, n_threads_ = 1;
    int64_t bytes_ = 0;
    std::vector<std::vector<char>> chunks_;
};

}  // namespace itr
