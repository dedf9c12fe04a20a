//! Raw declarations of include/ntgpu.h (ABI version 3), written by hand so that the crate builds without bindgen.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int};

pub const NTG_ABI_VERSION: c_int = 3;
pub const NTG_OK: c_int = 0;
pub const NTG_EIO: c_int = 1;
pub const NTG_EUNKNOWN_FORMAT: c_int = 2;
pub const NTG_EINVALID_START: c_int = 3;
pub const NTG_EINVALID_SEPARATOR: c_int = 4;
pub const NTG_EUNEQUAL_LENGTHS: c_int = 5;
pub const NTG_EUNEXPECTED_END: c_int = 6;
pub const NTG_EEMPTY_FILE: c_int = 7;
pub const NTG_EINVAL: c_int = 16;
pub const NTG_FMT_NONE: c_int = 0;
pub const NTG_FMT_FASTA: c_int = 1;
pub const NTG_FMT_FASTQ: c_int = 2;

#[repr(C)]
pub struct ntg_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct ntg_stream { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ntg_record {
    pub start: u64, pub id_b: u64, pub id_e: u64, pub seq_b: u64, pub seq_e: u64,
    pub qual_b: u64, pub qual_e: u64, pub all_e: u64, pub num_bases: u64, pub line: u64,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ntg_parse_error {
    pub kind: i32, pub format: i32, pub line: u64, pub record_index: u64, pub has_id: i32, pub id: [c_char; 236],
}
#[repr(C)]
pub struct ntg_records {
    pub format: i32, pub line_ending: i32, pub n_records: u64, pub records: *const ntg_record,
    pub error: ntg_parse_error, pub final_line: u64, pub final_byte: u64, pub _priv: *mut std::ffi::c_void,
}
#[repr(C)]
pub struct ntg_items {
    pub n_seqs: u64, pub n_items: u64, pub item_offs: *const u64, pub pos: *const u32, pub was_rc: *const u8,
    pub val_lo: *const u64, pub val_hi: *const u64, pub _priv: *mut std::ffi::c_void,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ntg_tally_config {
    pub k: u32, pub m: u32, pub allow_iupac: u32, pub has_query: u32, pub query: [u8; 64], pub flags: u32, pub qmask_score: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ntg_tallies {
    pub n_records: u64, pub n_bases: u64, pub n_kmers: u64, pub n_not_rc: u64, pub kmer_sum_lo: u64, pub kmer_sum_hi: u64,
    pub n_query: u64, pub n_minimizers: u64, pub minimizer_sum: u64, pub reserved: [u64; 7],
}

extern "C" {
    pub fn ntg_abi_version() -> c_int;
    pub fn ntg_create(device: c_int, out: *mut *mut ntg_ctx) -> c_int;
    pub fn ntg_destroy(ctx: *mut ntg_ctx);
    pub fn ntg_last_error(ctx: *const ntg_ctx) -> *const c_char;
    pub fn ntg_release_scratch(ctx: *mut ntg_ctx) -> c_int;
    pub fn ntg_alloc_pinned(bytes: usize, out: *mut *mut core::ffi::c_void) -> c_int;
    pub fn ntg_free_pinned(p: *mut core::ffi::c_void) -> c_int;
    pub fn ntg_parse_fastx_chunk(ctx: *mut ntg_ctx, bytes: *const u8, n: usize, format: c_int, at_eof: c_int,
                                 out: *mut *mut ntg_records, consumed: *mut u64) -> c_int;
    pub fn ntg_records_free(r: *mut ntg_records);
    pub fn ntg_normalize(ctx: *mut ntg_ctx, seqs: *const u8, offs: *const u64, n: usize, allow_iupac: c_int,
                         out: *mut u8, out_offs: *mut u64, changed: *mut u8) -> c_int;
    pub fn ntg_strip_returns(ctx: *mut ntg_ctx, seqs: *const u8, offs: *const u64, n: usize,
                             out: *mut u8, out_offs: *mut u64, changed: *mut u8) -> c_int;
    pub fn ntg_reverse_complement(ctx: *mut ntg_ctx, seqs: *const u8, offs: *const u64, n: usize, out: *mut u8) -> c_int;
    pub fn ntg_quality_mask(ctx: *mut ntg_ctx, seqs: *const u8, quals: *const u8, offs: *const u64, qual_offs: *const u64,
                            n: usize, score: u8, out: *mut u8) -> c_int;
    pub fn ntg_items_free(it: *mut ntg_items);
    pub fn ntg_canonical_kmers(ctx: *mut ntg_ctx, seqs: *const u8, rc: *const u8, offs: *const u64, n: usize, k: u32,
                               out: *mut *mut ntg_items) -> c_int;
    pub fn ntg_kmers(ctx: *mut ntg_ctx, seqs: *const u8, offs: *const u64, n: usize, k: u32, out: *mut *mut ntg_items) -> c_int;
    pub fn ntg_bit_kmers(ctx: *mut ntg_ctx, seqs: *const u8, offs: *const u64, n: usize, k: u32, canonical: c_int,
                         out: *mut *mut ntg_items) -> c_int;
    pub fn ntg_bit_minimizers(ctx: *mut ntg_ctx, seqs: *const u8, offs: *const u64, n: usize, k: u32, m: u32,
                              out: *mut *mut ntg_items) -> c_int;
    pub fn ntg_bitkmer_reverse_complement(ctx: *mut ntg_ctx, inp: *const u64, n: usize, k: u32, out: *mut u64) -> c_int;
    pub fn ntg_bitkmer_canonical(ctx: *mut ntg_ctx, inp: *const u64, n: usize, k: u32, out: *mut u64, was_rc: *mut u8) -> c_int;
    pub fn ntg_bitkmer_minimizer(ctx: *mut ntg_ctx, inp: *const u64, n: usize, k: u32, m: u32, out: *mut u64) -> c_int;
    pub fn ntg_tally_fastx(ctx: *mut ntg_ctx, bytes: *const u8, n: usize, cfg: *const ntg_tally_config,
                           out: *mut ntg_tallies, err: *mut ntg_parse_error) -> c_int;
    pub fn ntg_tally_fastx_file(ctx: *mut ntg_ctx, path: *const c_char, cfg: *const ntg_tally_config, threads: c_int,
                                out: *mut ntg_tallies, err: *mut ntg_parse_error) -> c_int;
    pub fn ntg_stream_open(ctx: *mut ntg_ctx, cfg: *const ntg_tally_config, out: *mut *mut ntg_stream) -> c_int;
    pub fn ntg_stream_feed(s: *mut ntg_stream, bytes: *const u8, n: usize) -> c_int;
    pub fn ntg_stream_acquire(s: *mut ntg_stream, ptr: *mut *mut u8, avail: *mut usize) -> c_int;
    pub fn ntg_stream_commit(s: *mut ntg_stream, n: usize) -> c_int;
    pub fn ntg_stream_feed_gz(s: *mut ntg_stream, gz: *const u8, n: usize, threads: c_int) -> c_int;
    pub fn ntg_stream_finish(s: *mut ntg_stream, out: *mut ntg_tallies, err: *mut ntg_parse_error) -> c_int;
    pub fn ntg_stream_close(s: *mut ntg_stream);
}
