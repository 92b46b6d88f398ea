// pbwire.hpp -- protobuf wire format and loom's file framing, without a protobuf runtime.
//
// Wire format: a message is a sequence of (tag = field << 3 | wire type) + payload;
// wire types 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32.  loom's schema is
// proto2 with no [packed = true] (src/schema.proto), so repeated scalars are written one
// tag per element; the readers here accept the packed form too, as every protobuf
// parser does.
// Files: src/protobuf_stream.hpp:100-144 (InFile), :258-282 (OutFile); zlib's gz layer
// reads gzip and plain files alike.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

namespace pbw {

enum { VARINT = 0, FIXED64 = 1, BYTES = 2, FIXED32 = 5 };

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    Reader(const uint8_t *b, size_t n) : p(b), end(b + n) {}
    explicit Reader(const std::vector<uint8_t> &v) : p(v.data()), end(v.data() + v.size()) {}
    bool more() const { return ok && p < end; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) { ok = false; return 0; }
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    uint32_t fixed32() {
        if (end - p < 4) { ok = false; p = end; return 0; }
        uint32_t v;
        std::memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    uint64_t fixed64() {
        if (end - p < 8) { ok = false; p = end; return 0; }
        uint64_t v;
        std::memcpy(&v, p, 8);
        p += 8;
        return v;
    }
    float f32() { const uint32_t u = fixed32(); float f; std::memcpy(&f, &u, 4); return f; }
    double f64() { const uint64_t u = fixed64(); double f; std::memcpy(&f, &u, 8); return f; }
    Reader sub() {   // payload of a length-delimited field
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Reader(end, 0); }
        Reader r(p, (size_t)n);
        p += n;
        return r;
    }
    bool next(uint32_t &field, uint32_t &wt) {
        if (!more()) return false;
        const uint64_t tag = varint();
        field = (uint32_t)(tag >> 3);
        wt = (uint32_t)(tag & 7);
        return ok;
    }
    void skip(uint32_t wt) {
        switch (wt) {
        case VARINT: varint(); break;
        case FIXED64: fixed64(); break;
        case BYTES: sub(); break;
        case FIXED32: fixed32(); break;
        default: ok = false;   // groups are not used by loom
        }
    }
    // one occurrence of a repeated scalar field, packed or not
    template <class Out> void rep_varint(uint32_t wt, Out &out) {
        if (wt == BYTES) { Reader r = sub(); while (r.more()) out.push_back((typename Out::value_type)r.varint()); ok = ok && r.ok; }
        else if (wt == VARINT) out.push_back((typename Out::value_type)varint());
        else ok = false;
    }
    template <class Out> void rep_f32(uint32_t wt, Out &out) {
        if (wt == BYTES) { Reader r = sub(); while (r.more()) out.push_back(r.f32()); ok = ok && r.ok; }
        else if (wt == FIXED32) out.push_back(f32());
        else ok = false;
    }
};

struct Writer {
    std::vector<uint8_t> buf;
    void clear() { buf.clear(); }
    void varint(uint64_t v) {
        while (v >= 0x80) { buf.push_back((uint8_t)(v | 0x80)); v >>= 7; }
        buf.push_back((uint8_t)v);
    }
    void tag(uint32_t field, uint32_t wt) { varint(((uint64_t)field << 3) | wt); }
    void u64(uint32_t field, uint64_t v) { tag(field, VARINT); varint(v); }
    void boolean(uint32_t field, bool v) { tag(field, VARINT); buf.push_back(v ? 1 : 0); }
    void f32(uint32_t field, float v) { tag(field, FIXED32); raw(&v, 4); }
    void f64(uint32_t field, double v) { tag(field, FIXED64); raw(&v, 8); }
    void msg(uint32_t field, const Writer &m) {
        tag(field, BYTES);
        varint(m.buf.size());
        buf.insert(buf.end(), m.buf.begin(), m.buf.end());
    }
    void raw(const void *p, size_t n) {
        const uint8_t *b = (const uint8_t *)p;
        buf.insert(buf.end(), b, b + n);
    }
};

inline bool ends_with(const std::string &s, const char *suffix) {
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

// protobuf::InFile (src/protobuf_stream.hpp:51-226)
class InFile {
public:
    explicit InFile(const char *path) : path_(path) {
        // plain stdin is read with read(2), exactly as many bytes as the next frame needs: a client that
        // pipes one request and waits for its response (loom.query's piped server, loom/runner.py:376-381)
        // must not be held up by zlib filling a buffer first
        if (path_ == "-") raw_fd_ = dup(STDIN_FILENO);
        else if (path_ == "-.gz") f_ = gzdopen(dup(STDIN_FILENO), "rb");
        else f_ = gzopen(path, "rb");
        if (f_) gzbuffer(f_, 1 << 20);
    }
    ~InFile() {
        if (f_) gzclose(f_);
        if (raw_fd_ >= 0) ::close(raw_fd_);
    }
    InFile(const InFile &) = delete;
    bool is_open() const { return f_ != nullptr || raw_fd_ >= 0; }
    uint64_t position() const { return position_; }
    // read(): the whole file is one message
    bool read_all(std::vector<uint8_t> &out) {
        out.clear();
        uint8_t chunk[1 << 16];
        for (;;) {
            const int n = rd(chunk, sizeof(chunk));
            if (n < 0) return false;
            if (n == 0) return true;
            out.insert(out.end(), chunk, chunk + n);
        }
    }
    // try_read_stream(): false at a clean end of stream; `bad` reports a truncated frame
    bool try_read_stream(std::vector<uint8_t> &msg, bool &bad) {
        uint8_t head[4];
        const int n = rd(head, 4);
        bad = false;
        if (n == 0) return false;
        if (n != 4) { bad = true; return false; }
        const uint32_t size = (uint32_t)head[0] | (uint32_t)head[1] << 8 | (uint32_t)head[2] << 16 | (uint32_t)head[3] << 24;
        msg.resize(size);
        if (size && rd(msg.data(), size) != (int)size) { bad = true; return false; }
        ++position_;
        return true;
    }
    // append the next frame's payload to `pool`; returns its size through `size`
    bool try_read_stream_into(std::vector<uint8_t> &pool, uint32_t &size, bool &bad) {
        uint8_t head[4];
        const int n = rd(head, 4);
        bad = false;
        if (n == 0) return false;
        if (n != 4) { bad = true; return false; }
        size = (uint32_t)head[0] | (uint32_t)head[1] << 8 | (uint32_t)head[2] << 16 | (uint32_t)head[3] << 24;
        const size_t at = pool.size();
        pool.resize(at + size);
        if (size && rd(pool.data() + at, size) != (int)size) { bad = true; return false; }
        ++position_;
        return true;
    }
    bool skip_stream(uint32_t &size, bool &bad) {
        uint8_t head[4];
        const int n = rd(head, 4);
        bad = false;
        if (n == 0) return false;
        if (n != 4) { bad = true; return false; }
        size = (uint32_t)head[0] | (uint32_t)head[1] << 8 | (uint32_t)head[2] << 16 | (uint32_t)head[3] << 24;
        uint8_t discard[1 << 14];       // read and drop: gzseek is very slow on uncompressed files
        for (uint32_t left = size; left;) {
            const unsigned take = left < sizeof(discard) ? left : (unsigned)sizeof(discard);
            if (rd(discard, take) != (int)take) { bad = true; return false; }
            left -= take;
        }
        ++position_;
        return true;
    }

private:
    int rd(void *buf, unsigned n) {
        if (raw_fd_ < 0) return gzread(f_, buf, n);
        unsigned have = 0;
        while (have < n) {
            const ssize_t k = ::read(raw_fd_, (char *)buf + have, n - have);
            if (k < 0) return -1;
            if (k == 0) break;
            have += (unsigned)k;
        }
        return (int)have;
    }
    std::string path_;
    gzFile f_ = nullptr;
    int raw_fd_ = -1;
    uint64_t position_ = 0;
};

// protobuf::OutFile (src/protobuf_stream.hpp:229-318)
class OutFile {
public:
    OutFile(const char *path, bool append = false) : path_(path) {
        const bool gz = ends_with(path_, ".gz");
        const char *mode = gz ? (append ? "ab" : "wb") : (append ? "abT" : "wbT");   // T = no compression
        if (path_ == "-" || path_ == "-.gz") f_ = gzdopen(dup(STDOUT_FILENO), mode);
        else f_ = gzopen(path, mode);
    }
    ~OutFile() { close(); }
    OutFile(const OutFile &) = delete;
    bool is_open() const { return f_ != nullptr; }
    bool write(const std::vector<uint8_t> &msg) {
        return msg.empty() || gzwrite(f_, msg.data(), (unsigned)msg.size()) == (int)msg.size();
    }
    bool write_stream(const std::vector<uint8_t> &msg) {
        const uint32_t n = (uint32_t)msg.size();
        const uint8_t head[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        return gzwrite(f_, head, 4) == 4 && write(msg);
    }
    bool flush() { return gzflush(f_, Z_SYNC_FLUSH) == Z_OK; }
    bool close() {
        if (!f_) return true;
        const int rc = gzclose(f_);
        f_ = nullptr;
        return rc == Z_OK;
    }

private:
    std::string path_;
    gzFile f_ = nullptr;
};

}  // namespace pbw
