// deflate_b200.hpp -- C++ mirror of the `deflate` crate's encode API (image-rs/deflate-rs, crate `deflate` 1.0.0)
// over the C ABI of deflate_b200.h.  Header only; link with -ldeflate_b200.
//
//   deflate_b200::deflate_bytes / _conf / _zlib[_conf] / _gzip[_conf]   src/lib.rs:137-286
//   deflate_b200::DeflateEncoder / ZlibEncoder / GzEncoder              src/writer.rs:89-467
//   deflate_b200::Compression, CompressionOptions                       src/compression_options.rs:31-196
//
// Errors: the one-shot functions throw deflate_b200::Error where the reference panics (src/lib.rs:145,186,190,196);
// the encoders throw it where the reference returns io::Error.  Without a CUDA device every call fails with
// DFL_E_NODEVICE: there is no CPU fallback behind this header.
#ifndef DEFLATE_B200_HPP
#define DEFLATE_B200_HPP

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "deflate_b200.h"

namespace deflate_b200 {

class Error : public std::runtime_error {
public:
    Error(int status, const std::string& where)
        : std::runtime_error(where + ": " + dfl_strerror(status) + detail()), status_(status) {}
    int status() const { return status_; }

private:
    static std::string detail() {
        const char* d = dfl_last_cuda_error();
        return (d && *d) ? std::string(" [") + d + "]" : std::string();
    }
    int status_;
};

enum class Compression { Fast, Default, Best };            // compression_options.rs:31-42
enum class MatchingType : uint8_t { Greedy = 0, Lazy = 1 };   // lz77.rs:26-37

struct CompressionOptions {                                // compression_options.rs:78-120
    uint16_t max_hash_checks = 128;
    uint16_t lazy_if_less_than = 32;
    MatchingType matching_type = MatchingType::Lazy;

    CompressionOptions() = default;
    CompressionOptions(uint16_t checks, uint16_t lazy, MatchingType mt) : max_hash_checks(checks), lazy_if_less_than(lazy), matching_type(mt) {}
    CompressionOptions(Compression c) {                     // impl From<Compression>, :188-196
        *this = c == Compression::Fast ? fast() : (c == Compression::Best ? high() : CompressionOptions());
    }
    static CompressionOptions high() { return {1768, 128, MatchingType::Lazy}; }         // :126-133
    static CompressionOptions fast() { return {1, 0, MatchingType::Greedy}; }            // :141-148
    static CompressionOptions huffman_only() { return {0, 0, MatchingType::Greedy}; }    // :155-162
    static CompressionOptions rle() { return {0, 0, MatchingType::Lazy}; }               // :171-178
    dfl_options c() const { return dfl_options{max_hash_checks, lazy_if_less_than, static_cast<uint8_t>(matching_type), 0}; }
};

namespace detail {
inline std::vector<uint8_t> oneshot(const uint8_t* in, size_t n, const CompressionOptions& o, int wrap,
                                    const std::vector<uint8_t>& gz_hdr, const char* what) {
    const dfl_options opt = o.c();
    std::vector<uint8_t> out(dfl_bound(n, wrap) + gz_hdr.size());
    size_t len = 0;
    const int rc = dfl_compress(in, n, &opt, wrap, gz_hdr.empty() ? nullptr : gz_hdr.data(), gz_hdr.size(), out.data(),
                                out.size(), &len);
    if (rc != DFL_OK) throw Error(rc, what);
    out.resize(len);
    return out;
}
}  // namespace detail

inline std::vector<uint8_t> deflate_bytes_conf(const uint8_t* in, size_t n, const CompressionOptions& o) {   // lib.rs:137
    return detail::oneshot(in, n, o, DFL_RAW, {}, "deflate_bytes_conf");
}
inline std::vector<uint8_t> deflate_bytes(const uint8_t* in, size_t n) { return deflate_bytes_conf(in, n, Compression::Default); }   // lib.rs:163
inline std::vector<uint8_t> deflate_bytes_zlib_conf(const uint8_t* in, size_t n, const CompressionOptions& o) {   // lib.rs:182
    return detail::oneshot(in, n, o, DFL_ZLIB, {}, "deflate_bytes_zlib_conf");
}
inline std::vector<uint8_t> deflate_bytes_zlib(const uint8_t* in, size_t n) { return deflate_bytes_zlib_conf(in, n, Compression::Default); }   // lib.rs:216
// gz_header: a complete RFC 1952 member header (what gzip_header::GzBuilder::into_header() returns, lib.rs:251);
// empty = the default header.  CRC-32 and ISIZE are computed on the device.
inline std::vector<uint8_t> deflate_bytes_gzip_conf(const uint8_t* in, size_t n, const CompressionOptions& o,
                                                    const std::vector<uint8_t>& gz_header = {}) {   // lib.rs:242
    return detail::oneshot(in, n, o, DFL_GZIP, gz_header, "deflate_bytes_gzip_conf");
}
inline std::vector<uint8_t> deflate_bytes_gzip(const uint8_t* in, size_t n) { return deflate_bytes_gzip_conf(in, n, Compression::Default); }   // lib.rs:284
/// Gives back the device scratch the library keeps between calls (dfl_trim); the next call allocates again.
inline void trim() { dfl_trim(); }

// write::{DeflateEncoder, ZlibEncoder, GzEncoder} (writer.rs:89-467).  The sink plays the role of `W: io::Write`:
// it is handed bytes and returns how many it took (0 = error, like io::ErrorKind::WriteZero).
class Encoder {
public:
    using Sink = std::function<size_t(const uint8_t*, size_t)>;
    Encoder(Sink sink, const CompressionOptions& o, int wrap, const std::vector<uint8_t>& gz_header = {})
        : sink_(std::move(sink)) {
        const dfl_options opt = o.c();
        h_ = dfl_encoder_new(&opt, wrap, gz_header.empty() ? nullptr : gz_header.data(), gz_header.size());
        if (!h_) throw Error(DFL_E_NODEVICE, "dfl_encoder_new");
    }
    Encoder(const Encoder&) = delete;
    Encoder& operator=(const Encoder&) = delete;
    ~Encoder() {   // writer.rs:139-152: finish on drop, errors ignored
        if (h_) {
            if (!finished_) { try { finish(); } catch (...) {} }
            dfl_encoder_free(h_);
        }
    }
    size_t write(const uint8_t* buf, size_t n) {            // Write::write, writer.rs:124-127
        size_t consumed = 0;
        check(dfl_encoder_write(h_, buf, n, &consumed), "dfl_encoder_write");
        drain();
        return consumed;
    }
    void write_all(const uint8_t* buf, size_t n) { while (n) { size_t c = write(buf, n); buf += c; n -= c; } }
    void flush() { check(dfl_encoder_flush(h_, DFL_FLUSH_SYNC), "dfl_encoder_flush"); drain(); }   // Flush::Sync, writer.rs:134-136
    void finish() { check(dfl_encoder_flush(h_, DFL_FLUSH_FINISH), "dfl_encoder_flush"); drain(); finished_ = true; }
    void reset(Sink next, const std::vector<uint8_t>& gz_header = {}) {   // writer.rs:112-116,403
        check(dfl_encoder_reset(h_, gz_header.empty() ? nullptr : gz_header.data(), gz_header.size()), "dfl_encoder_reset");
        drain();
        sink_ = std::move(next);
        finished_ = false;
    }
    uint32_t checksum() const { return dfl_encoder_checksum(h_); }   // writer.rs:248,429

private:
    void check(int rc, const char* where) { if (rc != DFL_OK) throw Error(rc, where); }
    void drain() {   // the inner.write loop of compress_until_done, writer.rs:15-58
        for (;;) {
            const uint8_t* p = nullptr;
            size_t n = 0;
            check(dfl_encoder_take_output(h_, &p, &n), "dfl_encoder_take_output");
            if (!n) return;
            const size_t took = sink_(p, n);
            if (!took) throw Error(DFL_E_STATE, "sink accepted no bytes");
            dfl_encoder_advance_output(h_, took);
        }
    }
    dfl_encoder* h_ = nullptr;
    Sink sink_;
    bool finished_ = false;
};
struct DeflateEncoder : Encoder { DeflateEncoder(Sink s, const CompressionOptions& o) : Encoder(std::move(s), o, DFL_RAW) {} };
struct ZlibEncoder : Encoder { ZlibEncoder(Sink s, const CompressionOptions& o) : Encoder(std::move(s), o, DFL_ZLIB) {} };
struct GzEncoder : Encoder {
    GzEncoder(Sink s, const CompressionOptions& o, const std::vector<uint8_t>& gz_header = {}) : Encoder(std::move(s), o, DFL_GZIP, gz_header) {}
};

}  // namespace deflate_b200
#endif  // DEFLATE_B200_HPP
