// Native multi-threaded reader for the reference's on-disk formats (include/orie_io.h).
//
// Replaces the per-file Python loop of lib/data.py:11-43 (load_data).  Host code only: plain C++17 + POSIX,
// compiled into liborie_io.so without any CUDA dependency.  A file the reader does not want to judge is handed
// back to the caller (fallback list) instead of being guessed at, so that every error the reference would
// raise is still raised by the reference-equivalent Python path (data.py).
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orie_io.h"

namespace {

thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

enum FileStatus { kMissing = 0, kParsed = 1, kFallback = 2, kIoError = 3 };

// whole file into `buf`; false if it is not a regular file (os.path.isfile, lib/data.py:23,27)
bool read_file(const std::string &path, std::vector<char> &buf, bool *io_error) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return false;
    const int fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
        *io_error = true;
        return true;
    }
    buf.resize((size_t)st.st_size);
    size_t got = 0;
    while (got < buf.size()) {
        const ssize_t n = read(fd, buf.data() + got, buf.size() - got);
        if (n < 0 && errno == EINTR) continue;
        if (n <= 0) break;
        got += (size_t)n;
    }
    close(fd);
    if (got != buf.size()) *io_error = true;
    return true;
}

// str.strip() / str.split() whitespace for ASCII text
inline bool py_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); }

// A token the fast path accepts: non-empty, only [0-9+-.eE], fully consumed by strtod.  Everything else
// (inf / nan spellings, underscores, hex floats, non-ASCII digits, doubled spaces ...) goes to the caller.
inline bool parse_token(const char *a, const char *b, double *out) {
    const size_t len = (size_t)(b - a);
    if (len == 0 || len > 63) return false;
    char tmp[64];
    for (size_t i = 0; i < len; ++i) {
        const char c = a[i];
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) return false;
        tmp[i] = c;
    }
    tmp[len] = 0;
    char *end = nullptr;
    *out = strtod(tmp, &end);          // glibc: correctly rounded, like Python's float()
    return end == tmp + len;
}

// rows of a text file (lib/data.py:24-26,31): every line stripped and split on single spaces; the table is cut
// to its shortest row; output columns = first five + last.
struct TextParser {
    std::vector<double> vals;     // all tokens of the file, row after row
    std::vector<int> widths;      // tokens per row

    FileStatus run(const std::vector<char> &buf, int need, std::vector<double> &out, int64_t *nrows) {
        vals.clear();
        widths.clear();
        *nrows = 0;
        if (buf.empty()) return kParsed;                              // empty file: the image has no rows
        const char *p = buf.data(), *end = p + buf.size();
        if (memchr(p, 0, buf.size())) return kFallback;
        for (const char *q = p; q < end; ++q)
            if ((unsigned char)*q >= 0x80) return kFallback;          // non-ASCII text: let Python decode it
        while (p < end) {
            // one line: up to '\n' (readlines() keeps universal newlines: '\r' alone also ends a line -> fallback)
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = nl ? nl : end;
            const char *a = p, *b = le;
            while (a < b && py_space(*a)) ++a;
            while (b > a && py_space(b[-1])) --b;
            for (const char *q = a; q < b; ++q)
                if (*q == '\r') return kFallback;
            int w = 0;
            const char *t = a;
            for (;;) {
                const char *sp = (const char *)memchr(t, ' ', (size_t)(b - t));
                const char *te = sp ? sp : b;
                double v;
                if (!parse_token(t, te, &v)) return kFallback;       // includes the empty token of a blank line
                vals.push_back(v);
                ++w;
                if (!sp) break;
                t = sp + 1;
            }
            widths.push_back(w);
            p = nl ? nl + 1 : end;
        }
        *nrows = (int64_t)widths.size();
        if (widths.empty()) return kParsed;                           // empty file: the image has no rows
        const int width = *std::min_element(widths.begin(), widths.end());
        if (width < need) return kFallback;                           // the Python path raises
        size_t at = 0;
        for (int w : widths) {
            const double *r = vals.data() + at;
            for (int c = 0; c < 5; ++c) out.push_back(r[c]);
            if (need == 6) out.push_back(r[width - 1]);
            at += (size_t)w;
        }
        return kParsed;
    }
};

// .npy written by np.save (torch_models/detect.py:102-105): C-ordered little-endian f4 / f8 matrix only
FileStatus parse_npy(const std::vector<char> &buf, int need, std::vector<double> &out, int64_t *nrows) {
    const unsigned char *p = (const unsigned char *)buf.data();
    const size_t n = buf.size();
    if (n < 10 || memcmp(p, "\x93NUMPY", 6) != 0) return kFallback;
    const int major = p[6];
    size_t hlen, hoff;
    if (major == 1) {
        hlen = (size_t)p[8] | ((size_t)p[9] << 8);
        hoff = 10;
    } else if (major == 2 || major == 3) {
        if (n < 12) return kFallback;
        hlen = (size_t)p[8] | ((size_t)p[9] << 8) | ((size_t)p[10] << 16) | ((size_t)p[11] << 24);
        hoff = 12;
    } else {
        return kFallback;
    }
    if (hoff + hlen > n) return kFallback;
    const std::string h((const char *)p + hoff, hlen);
    int item = 0;
    if (h.find("'descr': '<f8'") != std::string::npos) item = 8;
    else if (h.find("'descr': '<f4'") != std::string::npos) item = 4;
    else return kFallback;
    if (h.find("'fortran_order': False") == std::string::npos) return kFallback;
    const size_t sp = h.find("'shape': (");
    if (sp == std::string::npos) return kFallback;
    long long d0 = -1, d1 = -1;
    char tail = 0;
    if (sscanf(h.c_str() + sp + 10, "%lld, %lld%c", &d0, &d1, &tail) != 3 || tail != ')' || d0 < 0 || d1 < 0) return kFallback;
    const size_t body = hoff + hlen;
    if ((size_t)d0 * (size_t)d1 * (size_t)item != n - body) return kFallback;
    *nrows = d0;
    if (d0 == 0) return kParsed;
    if (d1 < need) return kFallback;
    const unsigned char *data = p + body;
    auto at = [&](long long r, long long c) -> double {
        const unsigned char *q = data + ((size_t)r * (size_t)d1 + (size_t)c) * (size_t)item;
        if (item == 8) { double v; memcpy(&v, q, 8); return v; }
        float v; memcpy(&v, q, 4); return (double)v;
    };
    for (long long r = 0; r < d0; ++r) {
        for (int c = 0; c < 5; ++c) out.push_back(at(r, c));
        if (need == 6) out.push_back(at(r, d1 - 1));
    }
    return kParsed;
}

struct ImageRecord {
    int thread = -1;
    size_t at = 0;        // offset (in doubles) inside the thread's buffer
    int64_t rows = 0;
    FileStatus status = kMissing;
};

}  // namespace

struct orie_rows {
    int64_t count = 0;
    int cols = 0;
    std::vector<int64_t> off;
    std::vector<double> data;
    std::vector<int64_t> fallback;
};

extern "C" const char *orie_io_last_error(void) { return g_error; }

extern "C" int orie_io_read_rows(const char *dir, const char *const *names, int64_t count, int with_conf, int threads,
                                 orie_rows_t **out) {
    if (!out) {
        set_error("orie_io_read_rows: out is NULL");
        return ORIE_IO_EINVAL;
    }
    *out = nullptr;
    if (!dir || count < 0 || (count > 0 && !names)) {
        set_error("orie_io_read_rows: bad arguments");
        return ORIE_IO_EINVAL;
    }
    const int need = with_conf ? 6 : 5;
    int nthreads = threads;
    if (nthreads <= 0) {
        const long cpus = sysconf(_SC_NPROCESSORS_ONLN);
        nthreads = (int)std::min<long>(std::max<long>(cpus, 1), 64);
    }
    nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, (count + 63) / 64));

    std::vector<ImageRecord> rec((size_t)count);
    std::vector<std::vector<double>> bufs((size_t)nthreads);
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> io_failed{-1};
    const std::string base = std::string(dir) + "/";

    auto work = [&](int tid) {
        std::vector<char> file;
        TextParser text;
        std::vector<double> &mine = bufs[(size_t)tid];
        for (;;) {
            const int64_t i0 = next.fetch_add(16);
            if (i0 >= count) break;
            const int64_t i1 = std::min<int64_t>(i0 + 16, count);
            for (int64_t i = i0; i < i1; ++i) {
                ImageRecord &r = rec[(size_t)i];
                r.thread = tid;
                r.at = mine.size();
                const std::string stem = base + names[i];
                bool io_error = false;
                FileStatus st = kMissing;
                int64_t nrows = 0;
                if (read_file(stem + ".txt", file, &io_error)) {
                    st = io_error ? kIoError : text.run(file, need, mine, &nrows);
                } else if (read_file(stem + ".npy", file, &io_error)) {
                    st = io_error ? kIoError : parse_npy(file, need, mine, &nrows);
                }
                if (st == kIoError) io_failed.store(i);
                if (st != kParsed) {
                    mine.resize(r.at);
                    nrows = 0;
                }
                r.rows = nrows;
                r.status = st;
            }
        }
    };
    try {
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto &t : pool) t.join();
    } catch (const std::exception &e) {
        set_error("orie_io_read_rows: %s", e.what());
        return ORIE_IO_ENOMEM;
    }
    if (io_failed.load() >= 0) {
        set_error("orie_io_read_rows: could not read the file of image '%s' in %s", names[io_failed.load()], dir);
        return ORIE_IO_EIO;
    }

    orie_rows *res = nullptr;
    try {
        res = new orie_rows();
        res->count = count;
        res->cols = need;
        res->off.resize((size_t)count + 1);
        res->off[0] = 0;
        for (int64_t i = 0; i < count; ++i) {
            res->off[(size_t)i + 1] = res->off[(size_t)i] + rec[(size_t)i].rows;
            if (rec[(size_t)i].status == kFallback) res->fallback.push_back(i);
        }
        res->data.resize((size_t)res->off[(size_t)count] * (size_t)need);
    } catch (const std::exception &e) {
        delete res;
        set_error("orie_io_read_rows: %s", e.what());
        return ORIE_IO_ENOMEM;
    }
    // gather the per-thread buffers into image order (parallel copy)
    std::atomic<int64_t> nextc{0};
    auto gather = [&]() {
        for (;;) {
            const int64_t i0 = nextc.fetch_add(256);
            if (i0 >= count) break;
            const int64_t i1 = std::min<int64_t>(i0 + 256, count);
            for (int64_t i = i0; i < i1; ++i) {
                const ImageRecord &r = rec[(size_t)i];
                if (r.rows > 0)
                    memcpy(res->data.data() + (size_t)res->off[(size_t)i] * (size_t)need, bufs[(size_t)r.thread].data() + r.at,
                           (size_t)r.rows * (size_t)need * sizeof(double));
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(gather);
        gather();
        for (auto &t : pool) t.join();
    }
    *out = res;
    return ORIE_IO_OK;
}

extern "C" int64_t orie_io_num_images(const orie_rows_t *r) { return r ? r->count : 0; }
extern "C" int64_t orie_io_num_rows(const orie_rows_t *r) { return r ? r->off[(size_t)r->count] : 0; }
extern "C" int orie_io_num_cols(const orie_rows_t *r) { return r ? r->cols : 0; }
extern "C" const int64_t *orie_io_offsets(const orie_rows_t *r) { return r ? r->off.data() : nullptr; }
extern "C" const double *orie_io_data(const orie_rows_t *r) { return r ? r->data.data() : nullptr; }
extern "C" int64_t orie_io_num_fallback(const orie_rows_t *r) { return r ? (int64_t)r->fallback.size() : 0; }
extern "C" const int64_t *orie_io_fallback(const orie_rows_t *r) { return r ? r->fallback.data() : nullptr; }
extern "C" void orie_io_free(orie_rows_t *r) { delete r; }
