// pcd_io.cu — PCD v0.7 I/O at scale, the step either side of the registration path:
// pcl::io::loadPCDFile (RealTimeRobot.cpp:34-35, scan_point.h:62) and pcl::io::savePCDFileASCII (RealTimeRobot.cpp:108-109,
// function.h:126-127) — plus DATA binary and DATA binary_compressed, which PCL writes for anything large.
//
//   read  : the whole file in one read; header parsed from the buffer; the data section decoded by all host threads
//           (ascii: the byte range is cut at line starts, lines counted, then parsed with std::from_chars — correctly
//           rounded like strtof; binary: strided gather; binary_compressed: LZF -> field-major arrays -> gather) straight
//           into (x, y, z, 1) records.  rtr_pcd_load decodes into the context's PINNED staging buffer and uploads from
//           there with one cudaMemcpyAsync.
//   write : ascii with 8 significant digits (what savePCDFileASCII prints), formatted by all host threads; binary;
//           binary_compressed (own LZF encoder, format-compatible with liblzf as vendored by PCL).
// Only x, y, z (float32) are kept, as the reference loads every file into pcl::PointCloud<pcl::PointXYZ>.
#include "common.cuh"
#include <algorithm>
#include <charconv>
#include <cstring>
#include <string>
#include <thread>

namespace {

struct PcdHeader {
    size_t npts = 0, data_off = 0;
    int mode = -1;                      // 0 ascii, 1 binary, 2 binary_compressed
    int col[3] = {-1, -1, -1};          // ascii: token index of x, y, z
    size_t off[3] = {0, 0, 0};          // binary: byte offset inside a record; compressed: offset of the field's array / npts
    size_t record = 0;                  // bytes per point
    int ntok = 0;
};

int fail(const char* what, const char* path) {
    fprintf(stderr, "rtr[pcd] %s: %s\n", what, path ? path : "");
    return RTR_ERR_INVALID;
}

int host_threads(size_t work_items) {
    unsigned hc = std::thread::hardware_concurrency();
    int t = (int)std::min<unsigned>(hc ? hc : 4u, 32u);
    if (work_items < 65536) t = 1;
    return std::max(t, 1);
}

template <class F>
void parallel_for(int nthreads, F&& fn) {
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads - 1);
    for (int t = 1; t < nthreads; ++t) th.emplace_back([&fn, t] { fn(t); });
    fn(0);
    for (auto& x : th) x.join();
}

// header lines end at the one starting with DATA; everything after its newline is the data section
int parse_header(const char* buf, size_t len, PcdHeader& h, const char* path) {
    std::vector<std::string> fields; std::vector<int> sizes, counts; std::vector<char> types;
    size_t w = 0, hgt = 1, pos = 0;
    bool have_points = false;
    while (pos < len) {
        size_t e = pos;
        while (e < len && buf[e] != '\n') ++e;
        std::string line(buf + pos, e - pos);
        pos = e + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> tok;
        size_t i = 0;
        while (i < line.size()) {
            while (i < line.size() && (line[i] == ' ' || line[i] == '\t')) ++i;
            size_t j = i;
            while (j < line.size() && line[j] != ' ' && line[j] != '\t') ++j;
            if (j > i) tok.emplace_back(line.substr(i, j - i));
            i = j;
        }
        if (tok.empty()) continue;
        const std::string& key = tok[0];
        if (key == "FIELDS" || key == "COLUMNS") fields.assign(tok.begin() + 1, tok.end());
        else if (key == "SIZE") { for (size_t k = 1; k < tok.size(); ++k) sizes.push_back(atoi(tok[k].c_str())); }
        else if (key == "TYPE") { for (size_t k = 1; k < tok.size(); ++k) types.push_back(tok[k][0]); }
        else if (key == "COUNT") { for (size_t k = 1; k < tok.size(); ++k) counts.push_back(atoi(tok[k].c_str())); }
        else if (key == "WIDTH" && tok.size() > 1) w = strtoull(tok[1].c_str(), nullptr, 10);
        else if (key == "HEIGHT" && tok.size() > 1) hgt = strtoull(tok[1].c_str(), nullptr, 10);
        else if (key == "POINTS" && tok.size() > 1) { h.npts = strtoull(tok[1].c_str(), nullptr, 10); have_points = true; }
        else if (key == "DATA" && tok.size() > 1) {
            if (tok[1] == "ascii") h.mode = 0; else if (tok[1] == "binary") h.mode = 1; else if (tok[1] == "binary_compressed") h.mode = 2;
            h.data_off = std::min(pos, len);
            break;
        }
    }
    if (h.mode < 0) return fail("no DATA line / unknown DATA mode", path);
    if (!have_points) h.npts = w * hgt;
    if (counts.empty()) counts.assign(fields.size(), 1);
    if (fields.empty() || fields.size() != sizes.size() || fields.size() != types.size() || fields.size() != counts.size())
        return fail("FIELDS / SIZE / TYPE / COUNT disagree", path);
    if (h.npts > (size_t)INT32_MAX) return fail("too many points", path);
    int c = 0; size_t o = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
        for (int a = 0; a < 3; ++a)
            if (fields[i].size() == 1 && fields[i][0] == "xyz"[a]) {
                if (types[i] != 'F' || sizes[i] != 4 || counts[i] != 1) return fail("x / y / z must be float32", path);
                h.col[a] = c; h.off[a] = o;
            }
        c += counts[i]; o += (size_t)sizes[i] * counts[i];
    }
    h.ntok = c; h.record = o;
    if (h.col[0] < 0 || h.col[1] < 0 || h.col[2] < 0) return fail("x / y / z field missing", path);
    return 0;
}

// ---------------------------------------------------------------- LZF (liblzf stream format, as used by PCL's binary_compressed)
// control byte c < 32: c + 1 literal bytes follow;  c >= 32: back reference, length (c >> 5) + 2 (if c >> 5 == 7 the next byte
// is added to the length), distance ((c & 31) << 8 | next byte) + 1; source and destination may overlap.
size_t lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
    size_t ip = 0, op = 0;
    while (ip < in_len) {
        unsigned ctrl = in[ip++];
        if (ctrl < 32) {
            size_t run = ctrl + 1;
            if (ip + run > in_len || op + run > out_len) return 0;
            memcpy(out + op, in + ip, run);
            ip += run; op += run;
        } else {
            size_t len = ctrl >> 5;
            if (len == 7) { if (ip >= in_len) return 0; len += in[ip++]; }
            if (ip >= in_len) return 0;
            size_t dist = ((size_t)(ctrl & 31) << 8 | in[ip++]) + 1;
            len += 2;
            if (dist > op || op + len > out_len) return 0;
            const unsigned char* ref = out + op - dist;
            if (dist >= len) memcpy(out + op, ref, len);
            else for (size_t k = 0; k < len; ++k) out[op + k] = ref[k];
            op += len;
        }
    }
    return op;
}

size_t lzf_compress(const unsigned char* in, size_t n, unsigned char* out, size_t cap) {
    const int HLOG = 16;
    std::vector<uint32_t> htab((size_t)1 << HLOG, 0u);      // position + 1 of the last occurrence of a 3-byte hash
    size_t ip = 0, op = 0;
    if (cap < n + n / 32 + 8) return 0;
    size_t ctrl_at = op++;                                   // control byte of the open literal run
    unsigned lit = 0;
    auto close_run = [&] { if (lit) out[ctrl_at] = (unsigned char)(lit - 1); else --op; };
    while (ip + 2 < n) {
        uint32_t v = ((uint32_t)in[ip] << 16) | ((uint32_t)in[ip + 1] << 8) | in[ip + 2];
        uint32_t hsh = ((v * 2654435761u) >> (32 - HLOG));
        uint32_t ref1 = htab[hsh];
        htab[hsh] = (uint32_t)(ip + 1);
        if (ref1) {
            size_t r = ref1 - 1, dist = ip - r;
            if (dist >= 1 && dist <= 8192 && in[r] == in[ip] && in[r + 1] == in[ip + 1] && in[r + 2] == in[ip + 2]) {
                size_t maxlen = std::min<size_t>(n - ip, 264), len = 3;
                while (len < maxlen && in[r + len] == in[ip + len]) ++len;
                close_run();
                size_t l = len - 2, off = dist - 1;
                if (l < 7) out[op++] = (unsigned char)((off >> 8) + (l << 5));
                else { out[op++] = (unsigned char)((off >> 8) + (7u << 5)); out[op++] = (unsigned char)(l - 7); }
                out[op++] = (unsigned char)(off & 0xff);
                ctrl_at = op++; lit = 0;
                ip += len;
                continue;
            }
        }
        out[op++] = in[ip++];
        if (++lit == 32) { out[ctrl_at] = 31; ctrl_at = op++; lit = 0; }
    }
    while (ip < n) {
        out[op++] = in[ip++];
        if (++lit == 32) { out[ctrl_at] = 31; ctrl_at = op++; lit = 0; }
    }
    close_run();
    return op;
}

// ---------------------------------------------------------------- decode the data section into (x, y, z, 1) records
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

int decode_ascii(const char* d, size_t len, const PcdHeader& h, float* out, const char* path) {
    const int T = host_threads(h.npts);
    std::vector<size_t> cut(T + 1), lines(T + 1, 0);
    cut[0] = 0; cut[T] = len;
    for (int t = 1; t < T; ++t) {
        size_t p = len / T * t;
        while (p < len && d[p] != '\n') ++p;
        cut[t] = std::min(p + 1, len);
    }
    auto nonblank = [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) if (!is_space(d[i])) return true; return false; };
    parallel_for(T, [&](int t) {
        size_t cnt = 0, p = cut[t];
        while (p < cut[t + 1]) {
            const char* nl = (const char*)memchr(d + p, '\n', cut[t + 1] - p);
            size_t e = nl ? (size_t)(nl - d) : cut[t + 1];
            if (nonblank(p, e)) ++cnt;
            p = e + 1;
        }
        lines[t + 1] = cnt;
    });
    for (int t = 0; t < T; ++t) lines[t + 1] += lines[t];
    if (lines[T] < h.npts) return fail("fewer data lines than POINTS", path);
    std::vector<int> bad(T, 0);
    parallel_for(T, [&](int t) {
        size_t row = lines[t], p = cut[t];
        while (p < cut[t + 1] && row < h.npts) {
            const char* nl = (const char*)memchr(d + p, '\n', cut[t + 1] - p);
            size_t e = nl ? (size_t)(nl - d) : cut[t + 1];
            if (nonblank(p, e)) {
                float* o = out + row * 4;
                o[0] = o[1] = o[2] = 0.f; o[3] = 1.f;
                int k = 0, got = 0;
                size_t i = p;
                while (i < e) {
                    while (i < e && is_space(d[i])) ++i;
                    size_t j = i;
                    while (j < e && !is_space(d[j])) ++j;
                    if (j > i) {
                        for (int a = 0; a < 3; ++a)
                            if (k == h.col[a]) {
                                const char* b = d + i;
                                if (*b == '+') ++b;
                                auto r = std::from_chars(b, d + j, o[a]);
                                if (r.ec == std::errc::result_out_of_range) o[a] = strtof(std::string(d + i, j - i).c_str(), nullptr);
                                else if (r.ec != std::errc()) bad[t] = 1;
                                ++got;
                            }
                        ++k;
                    }
                    i = j;
                }
                if (got != 3) bad[t] = 1;
                ++row;
            }
            p = e + 1;
        }
    });
    for (int t = 0; t < T; ++t) if (bad[t]) return fail("unparsable ascii row", path);
    return 0;
}

void gather_records(const unsigned char* d, const PcdHeader& h, size_t stride_pt, const size_t base[3], float* out) {
    const int T = host_threads(h.npts);
    parallel_for(T, [&](int t) {
        size_t b = h.npts * t / T, e = h.npts * (t + 1) / T;
        for (size_t i = b; i < e; ++i) {
            float* o = out + i * 4;
            for (int a = 0; a < 3; ++a) memcpy(o + a, d + base[a] + i * stride_pt, 4);
            o[3] = 1.f;
        }
    });
}

int decode(const std::vector<unsigned char>& file, const PcdHeader& h, float* out, const char* path) {
    const unsigned char* d = file.data() + h.data_off;
    size_t len = file.size() - h.data_off;
    if (h.npts == 0) return 0;
    if (h.mode == 0) return decode_ascii((const char*)d, len, h, out, path);
    if (h.mode == 1) {
        if (len < h.npts * h.record) return fail("binary data section too short", path);
        gather_records(d, h, h.record, h.off, out);
        return 0;
    }
    if (len < 8) return fail("compressed data section too short", path);
    uint32_t csize, usize;
    memcpy(&csize, d, 4); memcpy(&usize, d + 4, 4);
    if ((size_t)csize + 8 > len || (size_t)usize < h.npts * h.record) return fail("compressed sizes inconsistent", path);
    std::vector<unsigned char> raw(usize);
    if (lzf_decompress(d + 8, csize, raw.data(), usize) != usize) return fail("LZF stream corrupt", path);
    size_t base[3];
    for (int a = 0; a < 3; ++a) base[a] = h.off[a] * h.npts;      // field-major: every field's values are contiguous
    gather_records(raw.data(), h, 4, base, out);
    return 0;
}

int read_file(const char* path, std::vector<unsigned char>& file) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail("cannot open", path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return fail("cannot size", path); }
    file.resize((size_t)sz);
    size_t got = sz ? fread(file.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return fail("short read", path);
    return 0;
}

std::string header_text(size_t n, const char* mode) {
    char b[512];
    snprintf(b, sizeof b, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
             "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA %s\n", n, n, mode);
    return b;
}

}  // namespace

extern "C" {

int rtr_pcd_info(const char* path, int* n_points, int* data_mode) {
    if (!path) return fail("null path", nullptr);
    FILE* f = fopen(path, "rb");
    if (!f) return fail("cannot open", path);
    std::vector<unsigned char> head(4096);
    size_t got = fread(head.data(), 1, head.size(), f);
    fclose(f);
    PcdHeader h;
    if (int e = parse_header((const char*)head.data(), got, h, path)) return e;
    if (n_points) *n_points = (int)h.npts;
    if (data_mode) *data_mode = h.mode;
    return 0;
}

int rtr_pcd_read(const char* path, float* host_xyz1, int capacity, int* n_points) {
    if (!path || !n_points || capacity < 0 || (capacity > 0 && !host_xyz1)) return fail("bad argument", path);
    std::vector<unsigned char> file;
    if (int e = read_file(path, file)) return e;
    PcdHeader h;
    if (int e = parse_header((const char*)file.data(), file.size(), h, path)) return e;
    *n_points = (int)h.npts;
    if ((size_t)capacity < h.npts) return rtr_fail("pcd", "point buffer too small", RTR_ERR_CAPACITY);
    return decode(file, h, host_xyz1, path);
}

int rtr_pcd_load(rtr_context* ctx, const char* path, rtr_cloud** out) {
    if (!ctx || !path || !out) return fail("bad argument", path);
    std::vector<unsigned char> file;
    if (int e = read_file(path, file)) return e;
    PcdHeader h;
    if (int e = parse_header((const char*)file.data(), file.size(), h, path)) return e;
    RTR_CHECK(cudaSetDevice(ctx->device), "pcd");
    // decode straight into pinned staging (grown geometrically, owned by the context), one async copy from there
    size_t need = std::max<size_t>(h.npts, 1) * 16;
    if (need > ctx->io_pinned_cap) {
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "pcd");
        if (ctx->io_pinned) cudaFreeHost(ctx->io_pinned);
        ctx->io_pinned = nullptr; ctx->io_pinned_cap = 0;
        size_t cap = std::max(need, (size_t)1 << 20);
        cap += cap / 2;
        RTR_CHECK(cudaMallocHost(&ctx->io_pinned, cap), "pcd.pinned");
        ctx->io_pinned_cap = cap;
    } else {
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "pcd");       // a previous upload may still be reading the staging buffer
    }
    if (int e = decode(file, h, (float*)ctx->io_pinned, path)) return e;
    return rtr_cloud_upload(ctx, (const float*)ctx->io_pinned, (int)h.npts, out);
}

int rtr_pcd_write(const char* path, const float* host_xyz1, int n, int mode) {
    if (!path || n < 0 || (n > 0 && !host_xyz1) || mode < 0 || mode > 2) return fail("bad argument", path);
    FILE* f = fopen(path, "wb");
    if (!f) return fail("cannot create", path);
    const char* names[3] = {"ascii", "binary", "binary_compressed"};
    std::string hd = header_text((size_t)n, names[mode]);
    bool ok = fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    const size_t np = (size_t)n;
    if (mode == 0) {
        const int T = host_threads(np);
        std::vector<std::string> part(T);
        parallel_for(T, [&](int t) {
            size_t b = np * t / T, e = np * (t + 1) / T;
            std::string& s = part[t];
            s.reserve((e - b) * 36);
            char buf[64];
            for (size_t i = b; i < e; ++i) {
                char* p = buf;
                for (int a = 0; a < 3; ++a) {
                    float v = host_xyz1[i * 4 + a];
                    if (v != v) { memcpy(p, "nan", 3); p += 3; }                     // PCL prints "nan" for NaN fields
                    else { auto r = std::to_chars(p, buf + sizeof buf - 2, v, std::chars_format::general, 8); p = r.ptr; }
                    *p++ = (a == 2) ? '\n' : ' ';
                }
                s.append(buf, p - buf);
            }
        });
        for (int t = 0; t < T && ok; ++t) ok = fwrite(part[t].data(), 1, part[t].size(), f) == part[t].size();
    } else if (mode == 1) {
        std::vector<float> rec(np * 3);
        for (size_t i = 0; i < np; ++i) for (int a = 0; a < 3; ++a) rec[i * 3 + a] = host_xyz1[i * 4 + a];
        ok = ok && fwrite(rec.data(), 4, rec.size(), f) == rec.size();
    } else {
        std::vector<float> soa(np * 3);
        for (int a = 0; a < 3; ++a) for (size_t i = 0; i < np; ++i) soa[(size_t)a * np + i] = host_xyz1[i * 4 + a];
        size_t usize = soa.size() * 4, cap = usize + usize / 16 + 64;
        if (usize > 0xffffffffull) { fclose(f); return fail("cloud too large for binary_compressed", path); }
        std::vector<unsigned char> comp(cap);
        size_t csize = lzf_compress((const unsigned char*)soa.data(), usize, comp.data(), cap);
        if (usize && !csize) { fclose(f); return fail("LZF encoder overflow", path); }
        uint32_t cs = (uint32_t)csize, us = (uint32_t)usize;
        ok = ok && fwrite(&cs, 4, 1, f) == 1 && fwrite(&us, 4, 1, f) == 1 && fwrite(comp.data(), 1, csize, f) == csize;
    }
    if (fclose(f) != 0) ok = false;
    return ok ? 0 : fail("write failed", path);
}

int rtr_cloud_save(rtr_cloud* c, const char* path, int mode) {
    if (!c || !path) return fail("bad argument", path);
    std::vector<float> host((size_t)std::max(c->n, 1) * 4);
    if (int e = rtr_cloud_download(c, host.data())) return e;
    return rtr_pcd_write(path, host.data(), c->n, mode);
}

}  // extern "C"
