//! The `deflate` crate's public encode API (image-rs/deflate-rs, crate `deflate` 1.0.0) over
//! `libdeflate_b200.so` -- hand-written sm_100a CUDA kernels behind the C ABI of `include/deflate_b200.h`.
//!
//! Same names, argument meaning and error behaviour as the reference:
//! `deflate_bytes[_conf]`, `deflate_bytes_zlib[_conf]`, `deflate_bytes_gzip[_conf]` (`src/lib.rs:137-286`),
//! `write::{DeflateEncoder, ZlibEncoder, GzEncoder}` (`src/writer.rs:89-467`), `Compression`,
//! `CompressionOptions`, `MatchingType`, `SpecialOptions` (`src/compression_options.rs:31-196`).
//!
//! This file is the only `unsafe` a maintainer would add; it is NOT compiled in this repository's image (no
//! rustc there).  The C ABI it binds is exercised by the Python mirror, the C example and the C++ header.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::io::{self, Write};
use std::os::raw::{c_char, c_void};

#[cfg(feature = "gzip")]
pub use gzip_header::GzBuilder;

// ------------------------------------------------------------------------------------------------- C ABI
pub mod ffi {
    use super::*;

    #[repr(C)]
    #[derive(Clone, Copy, Debug)]
    pub struct dfl_options {
        pub max_hash_checks: u16,
        pub lazy_if_less_than: u16,
        pub matching_type: u8, // 0 = Greedy, 1 = Lazy
        pub special: u8,       // 0 = Normal
    }
    #[repr(C)]
    pub struct dfl_encoder {
        _private: [u8; 0],
    }
    #[repr(C)]
    pub struct dfl_comm {
        _private: [u8; 0],
    }

    pub const DFL_RAW: i32 = 0;
    pub const DFL_ZLIB: i32 = 1;
    pub const DFL_GZIP: i32 = 2;
    pub const DFL_FLUSH_SYNC: i32 = 1;
    pub const DFL_FLUSH_FINISH: i32 = 2;
    pub const DFL_OK: i32 = 0;
    pub const DFL_AGAIN: i32 = 1;
    pub const DFL_COMM_ID_BYTES: usize = 128;

    #[link(name = "deflate_b200")]
    extern "C" {
        pub fn dfl_strerror(status: i32) -> *const c_char;
        pub fn dfl_last_cuda_error() -> *const c_char;
        pub fn dfl_device_count() -> i32;
        pub fn dfl_bound(n: usize, wrap: i32) -> usize;
        pub fn dfl_trim() -> i32;
        pub fn dfl_compress(input: *const u8, n: usize, opt: *const dfl_options, wrap: i32, gz_hdr: *const u8,
                            gz_hdr_len: usize, out: *mut u8, out_cap: usize, out_len: *mut usize) -> i32;
        pub fn dfl_compress_batch(count: usize, input: *const *const u8, n: *const usize, opt: *const dfl_options,
                                  wrap: i32, out: *const *mut u8, out_cap: *const usize, out_len: *mut usize,
                                  status: *mut i32) -> i32;
        // device buffers (`*const c_void` = CUdeviceptr, `stream` = cudaStream_t)
        pub fn dfl_compress_device(d_in: *const c_void, n: usize, opt: *const dfl_options, wrap: i32, gz_hdr: *const u8,
                                   gz_hdr_len: usize, d_out: *mut c_void, out_cap: usize, out_len: *mut usize,
                                   stream: *mut c_void) -> i32;
        pub fn dfl_compress_device_piece(d_in: *const c_void, n_total: usize, dict_len: usize, opt: *const dfl_options,
                                         flush_mode: i32, d_out: *mut c_void, out_cap: usize, out_len: *mut usize,
                                         stream: *mut c_void) -> i32;
        pub fn dfl_compress_device_batch(count: usize, d_in: *const *const c_void, n: *const usize,
                                         opt: *const dfl_options, wrap: i32, d_out: *const *mut c_void,
                                         out_cap: *const usize, out_len: *mut usize, status: *mut i32) -> i32;
        // streaming handle
        pub fn dfl_encoder_new(opt: *const dfl_options, wrap: i32, gz_hdr: *const u8, gz_hdr_len: usize) -> *mut dfl_encoder;
        pub fn dfl_encoder_write(e: *mut dfl_encoder, buf: *const u8, n: usize, consumed: *mut usize) -> i32;
        pub fn dfl_encoder_flush(e: *mut dfl_encoder, mode: i32) -> i32;
        pub fn dfl_encoder_set_piece_bytes(e: *mut dfl_encoder, bytes: usize) -> i32;
        pub fn dfl_encoder_take_output(e: *mut dfl_encoder, p: *mut *const u8, len: *mut usize) -> i32;
        pub fn dfl_encoder_advance_output(e: *mut dfl_encoder, n: usize);
        pub fn dfl_encoder_checksum(e: *mut dfl_encoder) -> u32;
        pub fn dfl_encoder_reset(e: *mut dfl_encoder, gz_hdr: *const u8, gz_hdr_len: usize) -> i32;
        pub fn dfl_encoder_free(e: *mut dfl_encoder);
        // multi-GPU gather of compressed streams (NCCL)
        pub fn dfl_comm_unique_id(id: *mut u8) -> i32;
        pub fn dfl_comm_init(comm: *mut *mut dfl_comm, world: i32, rank: i32, id: *const u8) -> i32;
        pub fn dfl_comm_free(comm: *mut dfl_comm);
        pub fn dfl_gather_device(comm: *mut dfl_comm, d_src: *const c_void, n: usize, d_dst: *mut c_void, dst_cap: usize,
                                 sizes: *mut usize, root: i32, stream: *mut c_void) -> i32;
    }
}
use ffi::*;

// ------------------------------------------------------------------------------------------------- options
/// `src/lz77.rs:26-37`
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum MatchingType {
    Greedy,
    Lazy,
}
/// `src/compression_options.rs:52-59` (the other variants are unimplemented placeholders upstream)
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum SpecialOptions {
    Normal,
}
/// `src/compression_options.rs:31-42`
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum Compression {
    Fast,
    Default,
    Best,
}
impl std::default::Default for Compression {
    fn default() -> Compression {
        Compression::Default
    }
}
/// `src/compression_options.rs:78-120`
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub struct CompressionOptions {
    pub max_hash_checks: u16,
    pub lazy_if_less_than: u16,
    pub matching_type: MatchingType,
    pub special: SpecialOptions,
}
impl CompressionOptions {
    /// `src/compression_options.rs:126-133`
    pub const fn high() -> CompressionOptions {
        CompressionOptions { max_hash_checks: 1768, lazy_if_less_than: 128, matching_type: MatchingType::Lazy, special: SpecialOptions::Normal }
    }
    /// `src/compression_options.rs:141-148`
    pub const fn fast() -> CompressionOptions {
        CompressionOptions { max_hash_checks: 1, lazy_if_less_than: 0, matching_type: MatchingType::Greedy, special: SpecialOptions::Normal }
    }
    /// `src/compression_options.rs:155-162`
    pub const fn huffman_only() -> CompressionOptions {
        CompressionOptions { max_hash_checks: 0, lazy_if_less_than: 0, matching_type: MatchingType::Greedy, special: SpecialOptions::Normal }
    }
    /// `src/compression_options.rs:171-178`
    pub const fn rle() -> CompressionOptions {
        CompressionOptions { max_hash_checks: 0, lazy_if_less_than: 0, matching_type: MatchingType::Lazy, special: SpecialOptions::Normal }
    }
}
impl std::default::Default for CompressionOptions {
    /// `src/compression_options.rs:19-20,67-72`
    fn default() -> CompressionOptions {
        CompressionOptions { max_hash_checks: 128, lazy_if_less_than: 32, matching_type: MatchingType::Lazy, special: SpecialOptions::Normal }
    }
}
impl From<Compression> for CompressionOptions {
    /// `src/compression_options.rs:188-196`
    fn from(c: Compression) -> CompressionOptions {
        match c {
            Compression::Fast => CompressionOptions::fast(),
            Compression::Default => CompressionOptions::default(),
            Compression::Best => CompressionOptions::high(),
        }
    }
}
impl From<CompressionOptions> for dfl_options {
    fn from(o: CompressionOptions) -> dfl_options {
        dfl_options {
            max_hash_checks: o.max_hash_checks,
            lazy_if_less_than: o.lazy_if_less_than,
            matching_type: match o.matching_type { MatchingType::Greedy => 0, MatchingType::Lazy => 1 },
            special: 0,
        }
    }
}

fn status_text(status: i32) -> String {
    let base = unsafe { CStr::from_ptr(dfl_strerror(status)) }.to_string_lossy().into_owned();
    let detail = unsafe { CStr::from_ptr(dfl_last_cuda_error()) }.to_string_lossy().into_owned();
    if detail.is_empty() { base } else { format!("{} [{}]", base, detail) }
}
fn to_io(status: i32) -> io::Error {
    if status == DFL_AGAIN {
        io::ErrorKind::Interrupted.into() // "internal buffer full", src/compress.rs:117-120
    } else {
        io::Error::new(io::ErrorKind::Other, status_text(status))
    }
}

// ------------------------------------------------------------------------------------------------- one-shot
fn oneshot(input: &[u8], options: CompressionOptions, wrap: i32, gz_hdr: &[u8], what: &str) -> Vec<u8> {
    let opt: dfl_options = options.into();
    let cap = unsafe { dfl_bound(input.len(), wrap) } + gz_hdr.len();
    let mut out = Vec::<u8>::with_capacity(cap);
    let mut len = 0usize;
    let (hp, hl) = if gz_hdr.is_empty() { (std::ptr::null(), 0) } else { (gz_hdr.as_ptr(), gz_hdr.len()) };
    let rc = unsafe { dfl_compress(input.as_ptr(), input.len(), &opt, wrap, hp, hl, out.as_mut_ptr(), cap, &mut len) };
    // the reference `expect`s at the same places (src/lib.rs:145,186,190,196,255)
    assert!(rc == DFL_OK, "{}: {}", what, status_text(rc));
    unsafe { out.set_len(len) };
    out
}
/// `src/lib.rs:137`
pub fn deflate_bytes_conf<O: Into<CompressionOptions>>(input: &[u8], options: O) -> Vec<u8> {
    oneshot(input, options.into(), DFL_RAW, &[], "Write error!")
}
/// `src/lib.rs:163`
/// Not part of the crate's API: releases the device scratch the library keeps between calls for the calling
/// thread (and the pool of parked encoder resources); the next call allocates again.
pub fn trim_device_scratch() {
    unsafe { ffi::dfl_trim() };
}

pub fn deflate_bytes(input: &[u8]) -> Vec<u8> {
    deflate_bytes_conf(input, Compression::Default)
}
/// `src/lib.rs:182`
pub fn deflate_bytes_zlib_conf<O: Into<CompressionOptions>>(input: &[u8], options: O) -> Vec<u8> {
    oneshot(input, options.into(), DFL_ZLIB, &[], "Write error when writing compressed data!")
}
/// `src/lib.rs:216`
pub fn deflate_bytes_zlib(input: &[u8]) -> Vec<u8> {
    deflate_bytes_zlib_conf(input, Compression::Default)
}
/// `src/lib.rs:242`: the member header stays the gzip-header crate's business (`into_header()`, :251); CRC-32 and
/// ISIZE (:257-265) are computed on the device.
#[cfg(feature = "gzip")]
pub fn deflate_bytes_gzip_conf<O: Into<CompressionOptions>>(input: &[u8], options: O, gzip_header: GzBuilder) -> Vec<u8> {
    let hdr = gzip_header.into_header();
    oneshot(input, options.into(), DFL_GZIP, &hdr, "Write error when writing compressed data!")
}
/// `src/lib.rs:284`
#[cfg(feature = "gzip")]
pub fn deflate_bytes_gzip(input: &[u8]) -> Vec<u8> {
    deflate_bytes_gzip_conf(input, Compression::Default, GzBuilder::new())
}

// ------------------------------------------------------------------------------------------------- writers
pub mod write {
    use super::*;

    /// What the three encoders share: the handle, the sink, and the `inner.write` loop of `compress_until_done`
    /// (`src/writer.rs:15-58`) with its partial-write bookkeeping (`src/compress.rs:96-111,286-299`): the library
    /// lends bytes (`take_output`), the sink says how many it took (`advance_output`).
    struct Core<W: Write> {
        h: *mut dfl_encoder,
        inner: Option<W>,
    }
    impl<W: Write> Core<W> {
        fn new(writer: W, options: CompressionOptions, wrap: i32, gz_hdr: &[u8]) -> Core<W> {
            let opt: dfl_options = options.into();
            let (hp, hl) = if gz_hdr.is_empty() { (std::ptr::null(), 0) } else { (gz_hdr.as_ptr(), gz_hdr.len()) };
            let h = unsafe { dfl_encoder_new(&opt, wrap, hp, hl) };
            assert!(!h.is_null(), "dfl_encoder_new failed: {}", status_text(-4));
            Core { h, inner: Some(writer) }
        }
        fn drain(&mut self) -> io::Result<()> {
            loop {
                let (mut p, mut n) = (std::ptr::null(), 0usize);
                let rc = unsafe { dfl_encoder_take_output(self.h, &mut p, &mut n) };
                if rc != DFL_OK { return Err(to_io(rc)); }
                if n == 0 { return Ok(()); }
                let wrote = self.inner.as_mut().expect("Missing writer!").write(unsafe { std::slice::from_raw_parts(p, n) })?;
                if wrote == 0 { return Err(io::ErrorKind::WriteZero.into()); }
                unsafe { dfl_encoder_advance_output(self.h, wrote) };
            }
        }
        fn write(&mut self, buf: &[u8]) -> io::Result<usize> {
            let mut consumed = 0usize;
            let rc = unsafe { dfl_encoder_write(self.h, buf.as_ptr(), buf.len(), &mut consumed) };
            if rc != DFL_OK { return Err(to_io(rc)); }
            self.drain()?; // src/compress.rs:96-124: pending output goes to the sink before more is compressed
            Ok(consumed)
        }
        fn flush(&mut self, mode: i32) -> io::Result<()> {
            let rc = unsafe { dfl_encoder_flush(self.h, mode) };
            if rc != DFL_OK { return Err(to_io(rc)); }
            self.drain()
        }
        fn finish(&mut self) -> io::Result<W> {
            self.flush(DFL_FLUSH_FINISH)?;
            Ok(self.inner.take().expect("Missing writer!"))
        }
        fn reset(&mut self, w: W, gz_hdr: &[u8]) -> io::Result<W> {
            // src/writer.rs:112-116: the current stream is finished into the old sink, then the new one starts.
            // dfl_encoder_reset does the finishing; the finished stream's bytes stay available to take_output.
            let (hp, hl) = if gz_hdr.is_empty() { (std::ptr::null(), 0) } else { (gz_hdr.as_ptr(), gz_hdr.len()) };
            let rc = unsafe { dfl_encoder_reset(self.h, hp, hl) };
            if rc != DFL_OK { return Err(to_io(rc)); }
            self.drain()?;
            Ok(std::mem::replace(self.inner.as_mut().expect("Missing writer!"), w))
        }
        fn checksum(&self) -> u32 {
            unsafe { dfl_encoder_checksum(self.h) }
        }
    }
    impl<W: Write> Drop for Core<W> {
        /// `src/writer.rs:139-152,281-290,458-467`: finish on drop unless panicking; errors are ignored there too
        fn drop(&mut self) {
            if self.inner.is_some() && !std::thread::panicking() {
                let _ = self.flush(DFL_FLUSH_FINISH);
            }
            unsafe { dfl_encoder_free(self.h) };
        }
    }

    macro_rules! encoder {
        ($name:ident, $wrap:expr, $doc:expr) => {
            #[doc = $doc]
            pub struct $name<W: Write> {
                core: Core<W>,
            }
            impl<W: Write> $name<W> {
                pub fn new<O: Into<CompressionOptions>>(writer: W, options: O) -> $name<W> {
                    $name { core: Core::new(writer, options.into(), $wrap, &[]) }
                }
                /// Encode all pending data, write the trailer (if any) and hand the sink back.
                pub fn finish(mut self) -> io::Result<W> {
                    self.core.finish()
                }
                /// Finish the current stream, start a new one into `writer`, return the old sink.
                pub fn reset(&mut self, writer: W) -> io::Result<W> {
                    self.core.reset(writer, &[])
                }
            }
            impl<W: Write> Write for $name<W> {
                fn write(&mut self, buf: &[u8]) -> io::Result<usize> {
                    self.core.write(buf)
                }
                /// `Flush::Sync`: the block is closed and `00 00 FF FF` appended (`src/compress.rs:258-261`)
                fn flush(&mut self) -> io::Result<()> {
                    self.core.flush(DFL_FLUSH_SYNC)
                }
            }
        };
    }
    encoder!(DeflateEncoder, DFL_RAW, "`src/writer.rs:89-152`");
    encoder!(ZlibEncoder, DFL_ZLIB, "`src/writer.rs:183-290`");
    impl<W: Write> ZlibEncoder<W> {
        /// `src/writer.rs:248`: Adler-32 of the input consumed so far
        pub fn checksum(&self) -> u32 {
            self.core.checksum()
        }
    }

    /// `src/writer.rs:331-467`
    #[cfg(feature = "gzip")]
    pub struct GzEncoder<W: Write> {
        core: Core<W>,
    }
    #[cfg(feature = "gzip")]
    impl<W: Write> GzEncoder<W> {
        pub fn new<O: Into<CompressionOptions>>(writer: W, options: O) -> GzEncoder<W> {
            GzEncoder::from_builder(GzBuilder::new(), writer, options)
        }
        /// `src/writer.rs:346`
        pub fn from_builder<O: Into<CompressionOptions>>(builder: GzBuilder, writer: W, options: O) -> GzEncoder<W> {
            GzEncoder { core: Core::new(writer, options.into(), DFL_GZIP, &builder.into_header()) }
        }
        pub fn finish(mut self) -> io::Result<W> {
            self.core.finish()
        }
        pub fn reset(&mut self, writer: W) -> io::Result<W> {
            self.reset_with_builder(writer, GzBuilder::new())
        }
        /// `src/writer.rs:403`
        pub fn reset_with_builder(&mut self, writer: W, builder: GzBuilder) -> io::Result<W> {
            self.core.reset(writer, &builder.into_header())
        }
        /// `src/writer.rs:429`: CRC-32 of the input consumed so far
        pub fn checksum(&self) -> u32 {
            self.core.checksum()
        }
    }
    #[cfg(feature = "gzip")]
    impl<W: Write> Write for GzEncoder<W> {
        fn write(&mut self, buf: &[u8]) -> io::Result<usize> {
            self.core.write(buf)
        }
        fn flush(&mut self) -> io::Result<()> {
            self.core.flush(DFL_FLUSH_SYNC)
        }
    }
}
