// gpr_io.cu — host-only helpers behind the C-ABI: a minimal PCD v0.7 reader.
//
// The reference reads its object clouds with PCL (pcl::io::loadPCDFile, /root/reference/src/gp_node.cpp:557); a C++
// caller of the drop-in would otherwise still need PCL just to obtain gp_regression::Data.  Supported: DATA ascii /
// binary / binary_compressed (LZF, struct-of-arrays after decompression), 4-byte float fields x, y, z anywhere in the
// record (the reference's resources/*.pcd are float32 xyz[+rgba]); other fields are skipped.  No GPU work here.
#include "../../include/gpr_c_api.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace {

// LZF: control byte < 32 = literal run of ctrl+1 bytes; otherwise a back reference of length (ctrl >> 5) + 2
// (length field 7 = one more length byte) at distance ((ctrl & 0x1f) << 8 | next byte) + 1.
bool lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
    size_t i = 0, o = 0;
    while (i < in_len) {
        unsigned ctrl = in[i++];
        if (ctrl < 32) {
            const size_t run = ctrl + 1;
            if (i + run > in_len || o + run > out_len) return false;
            memcpy(out + o, in + i, run);
            i += run; o += run;
        } else {
            size_t len = ctrl >> 5;
            if (len == 7) { if (i >= in_len) return false; len += in[i++]; }
            if (i >= in_len) return false;
            const size_t dist = ((size_t)(ctrl & 0x1f) << 8) + in[i++] + 1;
            len += 2;
            if (dist > o || o + len > out_len) return false;
            for (size_t k = 0; k < len; ++k, ++o) out[o] = out[o - dist];       // may overlap: byte by byte
        }
    }
    return o == out_len;
}

struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; size_t offset = 0; };

}  // namespace

void gpr_set_last_error_internal(const char* msg);      // gpr_c_api.cu: the thread-local message behind gpr_last_error()

extern "C" {

void gpr_free(void* p) { free(p); }

int gpr_pcd_read_xyz(const char* path, double** x, double** y, double** z, size_t* n) {
    auto bad = [&](const std::string& msg) {
        gpr_set_last_error_internal((std::string(path ? path : "(null)") + ": " + msg).c_str());
        return (int)GPR_ERR_INVALID;
    };
    if (!path || !x || !y || !z || !n) return bad("null pointer");
    *x = *y = *z = nullptr; *n = 0;
    FILE* fh = fopen(path, "rb");
    if (!fh) return bad("cannot open");
    std::vector<unsigned char> raw;
    {
        unsigned char buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, fh)) > 0) raw.insert(raw.end(), buf, buf + got);
        fclose(fh);
    }
    // header: text lines up to and including "DATA <mode>"
    std::vector<Field> fields;
    size_t points = 0, width = 0, height = 1, pos = 0;
    bool have_points = false;
    std::string mode;
    while (pos < raw.size()) {
        size_t end = pos;
        while (end < raw.size() && raw[end] != '\n') ++end;
        if (end >= raw.size()) return bad("no DATA line");
        std::string line(raw.begin() + (std::ptrdiff_t)pos, raw.begin() + (std::ptrdiff_t)end);
        pos = end + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key, tok;
        ss >> key;
        std::vector<std::string> vals;
        while (ss >> tok) vals.push_back(tok);
        if (key == "FIELDS") { fields.resize(vals.size()); for (size_t i = 0; i < vals.size(); ++i) fields[i].name = vals[i]; }
        else if (key == "SIZE" || key == "TYPE" || key == "COUNT") {
            if (vals.size() != fields.size()) return bad("header field counts disagree");
            for (size_t i = 0; i < vals.size(); ++i) {
                if (key == "SIZE") fields[i].size = atoi(vals[i].c_str());
                else if (key == "TYPE") fields[i].type = vals[i].empty() ? 'F' : vals[i][0];
                else fields[i].count = atoi(vals[i].c_str());
            }
        }
        else if (key == "WIDTH" && !vals.empty()) width = strtoull(vals[0].c_str(), nullptr, 10);
        else if (key == "HEIGHT" && !vals.empty()) height = strtoull(vals[0].c_str(), nullptr, 10);
        else if (key == "POINTS" && !vals.empty()) { points = strtoull(vals[0].c_str(), nullptr, 10); have_points = true; }
        else if (key == "DATA") { if (vals.empty()) return bad("empty DATA line"); mode = vals[0]; break; }
    }
    if (mode.empty() || fields.empty()) return bad("not a PCD file (no FIELDS / DATA)");
    if (!have_points) points = width * height;
    if (points == 0 || points > ((size_t)1 << 28)) return bad("implausible POINTS");
    size_t rec = 0;
    int ix[3] = {-1, -1, -1};
    for (size_t i = 0; i < fields.size(); ++i) {
        if (fields[i].size <= 0 || fields[i].count <= 0 || fields[i].size > 8 || fields[i].count > 4096) return bad("bad SIZE / COUNT");
        fields[i].offset = rec;
        rec += (size_t)fields[i].size * fields[i].count;
        for (int c = 0; c < 3; ++c) if (fields[i].name == std::string(1, "xyz"[c])) ix[c] = (int)i;
    }
    for (int c = 0; c < 3; ++c) {
        if (ix[c] < 0) return bad("no x / y / z field");
        if (mode != "ascii" && (fields[ix[c]].size != 4 || fields[ix[c]].type != 'F')) return bad("x / y / z must be 4-byte floats");
    }
    double* out[3];
    for (int c = 0; c < 3; ++c) {
        out[c] = (double*)malloc(points * sizeof(double));
        if (!out[c]) { for (int d = 0; d < c; ++d) free(out[d]); return bad("out of host memory"); }
    }
    auto drop = [&](const std::string& msg) { for (int c = 0; c < 3; ++c) free(out[c]); return bad(msg); };
    if (mode == "ascii") {
        std::string text(raw.begin() + (std::ptrdiff_t)pos, raw.end());
        std::istringstream ss(text);
        size_t col_of[3] = {0, 0, 0}, ncol = 0;
        for (size_t i = 0; i < fields.size(); ++i) {
            for (int c = 0; c < 3; ++c) if ((int)i == ix[c]) col_of[c] = ncol;
            ncol += (size_t)fields[i].count;
        }
        std::string tok;
        for (size_t p = 0; p < points; ++p)
            for (size_t col = 0; col < ncol; ++col) {
                if (!(ss >> tok)) return drop("ascii data ends early");
                for (int c = 0; c < 3; ++c) if (col == col_of[c]) out[c][p] = (double)strtof(tok.c_str(), nullptr);   // float32 like the binary modes
            }
    } else if (mode == "binary") {
        if (raw.size() - pos < points * rec) return drop("binary data ends early");
        for (size_t p = 0; p < points; ++p)
            for (int c = 0; c < 3; ++c) {
                float v;
                memcpy(&v, raw.data() + pos + p * rec + fields[ix[c]].offset, 4);
                out[c][p] = (double)v;
            }
    } else if (mode == "binary_compressed") {
        if (raw.size() - pos < 8) return drop("compressed data ends early");
        unsigned int csize, usize;
        memcpy(&csize, raw.data() + pos, 4);
        memcpy(&usize, raw.data() + pos + 4, 4);
        if (raw.size() - pos - 8 < csize || (size_t)usize != points * rec) return drop("compressed sizes disagree with the header");
        std::vector<unsigned char> buf(usize);
        if (!lzf_decompress(raw.data() + pos + 8, csize, buf.data(), usize)) return drop("corrupt LZF stream");
        for (int c = 0; c < 3; ++c) {                                   // struct of arrays: all of field 0, then field 1, ...
            const unsigned char* src = buf.data() + fields[ix[c]].offset * points;
            for (size_t p = 0; p < points; ++p) { float v; memcpy(&v, src + 4 * p, 4); out[c][p] = (double)v; }
        }
    } else {
        return drop("unsupported DATA mode " + mode);
    }
    *x = out[0]; *y = out[1]; *z = out[2]; *n = points;
    return GPR_OK;
}

}  // extern "C"
