//! `needletail_b200` — needletail's public surface for the FASTX hot path, served by libntgpu (B200, sm_100a).
//!
//! Same names and meanings as the reference crate (`src/lib.rs:56-57`): [`parse_fastx_file`], [`parse_fastx_reader`],
//! [`parse_fastx_stdin`], the [`FastxReader`] trait, [`SequenceRecord`] and the [`Sequence`] methods.  The byte stream is read
//! (and gunzipped) on the host exactly where the reference does it (`src/parser/mod.rs:85-150`); every *computation* —
//! finding and validating records, normalize, reverse complement, canonical / bit k-mers, minimizers — is a call into the C ABI
//! of `include/ntgpu.h`.  There is no CPU fallback: without a B200 `Context::new` fails.
//!
//! SOURCE ONLY: the image this repository is built in has no Rust toolchain, so this crate has not been compiled or run here.
//! The ABI underneath is tested from C++ (`tests/cpp/test_host_mirror.cpp`) and Python (`tests/test_gpu_*.py`).
#[cfg(not(feature = "bindgen"))]
pub mod sys;
#[cfg(feature = "bindgen")]
#[allow(non_camel_case_types, non_upper_case_globals, dead_code)]
pub mod sys {
    include!(concat!(env!("OUT_DIR"), "/ntgpu.rs"));
}

use std::borrow::Cow;
use std::ffi::{CStr, CString};
use std::fmt;
use std::fs::File;
use std::io::{self, Read};
use std::path::Path;
use std::ptr;
use std::sync::{Arc, Mutex};

// ------------------------------------------------------------------------------------------------ errors (src/errors.rs)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Format { Fasta, Fastq }
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum LineEnding { Windows, Unix }
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ParseErrorKind { Io, UnknownFormat, InvalidStart, InvalidSeparator, UnequalLengths, UnexpectedEnd, EmptyFile }
#[derive(Clone, Debug, Default, PartialEq, Eq)]
pub struct ErrorPosition { pub line: u64, pub id: Option<String> }
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct ParseError { pub msg: String, pub kind: ParseErrorKind, pub position: ErrorPosition, pub format: Option<Format> }
impl fmt::Display for ParseError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "{:?}: {} (line {})", self.kind, self.msg, self.position.line) }
}
impl std::error::Error for ParseError {}
impl From<io::Error> for ParseError {
    fn from(e: io::Error) -> Self { ParseError { msg: e.to_string(), kind: ParseErrorKind::Io, position: ErrorPosition::default(), format: None } }
}
fn kind_of(code: i32) -> ParseErrorKind {
    match code {
        2 => ParseErrorKind::UnknownFormat, 3 => ParseErrorKind::InvalidStart, 4 => ParseErrorKind::InvalidSeparator,
        5 => ParseErrorKind::UnequalLengths, 6 => ParseErrorKind::UnexpectedEnd, 7 => ParseErrorKind::EmptyFile, _ => ParseErrorKind::Io,
    }
}
fn format_of(code: i32) -> Option<Format> { match code { 1 => Some(Format::Fasta), 2 => Some(Format::Fastq), _ => None } }
fn parse_error_from(e: &sys::ntg_parse_error, line_base: u64) -> ParseError {
    let id = if e.has_id != 0 { Some(unsafe { CStr::from_ptr(e.id.as_ptr()) }.to_string_lossy().into_owned()) } else { None };
    let kind = kind_of(e.kind);
    ParseError { msg: format!("{:?}", kind), kind, position: ErrorPosition { line: e.line + line_base, id }, format: format_of(e.format) }
}

// ------------------------------------------------------------------------------------------------ context
/// One CUDA device + its streams (`ntg_ctx`).  Not `Sync`: same contract as the reference's `&mut self` readers.
pub struct Context { raw: *mut sys::ntg_ctx }
unsafe impl Send for Context {}
impl Context {
    pub fn new(device: i32) -> Result<Self, ParseError> {
        let mut raw = ptr::null_mut();
        let st = unsafe { sys::ntg_create(device, &mut raw) };
        if st != sys::NTG_OK { return Err(lib_error(ptr::null(), st)); }
        Ok(Context { raw })
    }
    fn check(&self, st: i32) -> Result<(), ParseError> { if st == sys::NTG_OK { Ok(()) } else { Err(lib_error(self.raw, st)) } }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::ntg_destroy(self.raw) } } }
fn lib_error(ctx: *const sys::ntg_ctx, st: i32) -> ParseError {
    let msg = unsafe { CStr::from_ptr(sys::ntg_last_error(ctx)) }.to_string_lossy().into_owned();
    ParseError { msg: format!("libntgpu status {}: {}", st, msg), kind: ParseErrorKind::Io, position: ErrorPosition::default(), format: None }
}
/// The process-wide default context (device `NTGPU_DEVICE`, default 0), created on first use.
pub fn default_context() -> Result<Arc<Mutex<Context>>, ParseError> {
    static CTX: Mutex<Option<Arc<Mutex<Context>>>> = Mutex::new(None);
    let mut g = CTX.lock().unwrap();
    if g.is_none() {
        let dev = std::env::var("NTGPU_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
        *g = Some(Arc::new(Mutex::new(Context::new(dev)?)));
    }
    Ok(g.as_ref().unwrap().clone())
}

// ------------------------------------------------------------------------------------------------ records (src/parser/record.rs, utils.rs)
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct Position { line: u64, byte: u64 }
impl Position {
    pub fn new(line: u64, byte: u64) -> Self { Position { line, byte } }
    pub fn line(&self) -> u64 { self.line }
    pub fn byte(&self) -> u64 { self.byte }
}
/// A record borrowing the reader's window, like the reference's (`record.rs:21-55`): valid until the next `next()`.
pub struct SequenceRecord<'a> { buf: &'a [u8], row: sys::ntg_record, format: Format, position: Position, line_ending: LineEnding }
impl<'a> SequenceRecord<'a> {
    pub fn format(&self) -> Format { self.format }
    pub fn id(&self) -> &[u8] { &self.buf[self.row.id_b as usize..self.row.id_e as usize] }
    pub fn raw_seq(&self) -> &[u8] { &self.buf[self.row.seq_b as usize..self.row.seq_e as usize] }
    /// `raw_seq` without line breaks (`record.rs:84-95`): borrowed for single-line sequences
    pub fn seq(&self) -> Cow<'_, [u8]> {
        let raw = self.raw_seq();
        if raw.iter().any(|&b| b == b'\n' || b == b'\r') { Cow::Owned(raw.iter().copied().filter(|&b| b != b'\n' && b != b'\r').collect()) } else { Cow::Borrowed(raw) }
    }
    pub fn qual(&self) -> Option<&[u8]> { if self.format == Format::Fastq { Some(&self.buf[self.row.qual_b as usize..self.row.qual_e as usize]) } else { None } }
    pub fn all(&self) -> &[u8] { &self.buf[self.row.start as usize..self.row.all_e as usize] }
    pub fn num_bases(&self) -> usize { self.row.num_bases as usize }
    pub fn start_line_number(&self) -> u64 { self.position.line }
    pub fn position(&self) -> &Position { &self.position }
    pub fn line_ending(&self) -> LineEnding { self.line_ending }
}
impl<'a> Sequence<'a> for SequenceRecord<'a> {
    fn sequence(&'a self) -> &'a [u8] { self.raw_seq() }          // record.rs:181-185
}

pub trait FastxReader: Send {
    fn next(&mut self) -> Option<Result<SequenceRecord<'_>, ParseError>>;
    fn position(&self) -> &Position;
    fn line_ending(&self) -> Option<LineEnding>;
}

/// The incremental reader: a sliding window over the (decompressed) stream, scanned on the device window by window with
/// `ntg_parse_fastx_chunk` — the refill loop of `fastq.rs:312-384` / `fasta.rs:291-346` with the record search on the GPU.
struct GpuReader<R: Read + Send> {
    ctx: Arc<Mutex<Context>>,
    src: R,
    win: Vec<u8>,              // unconsumed bytes of the stream
    eof: bool,
    format: i32,
    recs: *mut sys::ntg_records,
    next_row: u64,
    consumed: u64,
    byte_base: u64, line_base: u64,
    position: Position,
    line_ending: Option<LineEnding>,
    done: bool,
}
unsafe impl<R: Read + Send> Send for GpuReader<R> {}
const WINDOW: usize = 256 << 20;
impl<R: Read + Send> GpuReader<R> {
    fn refill(&mut self) -> Result<(), ParseError> {
        // drop what the last scan consumed, top the window up, scan again
        if !self.recs.is_null() {
            let r = unsafe { &*self.recs };
            self.line_base += r.final_line.saturating_sub(1);
            unsafe { sys::ntg_records_free(self.recs) };
            self.recs = ptr::null_mut();
        }
        self.win.drain(..self.consumed as usize);
        self.byte_base += self.consumed;
        self.consumed = 0;
        let mut want = WINDOW.max(self.win.len() * 2);
        loop {
            while !self.eof && self.win.len() < want {
                let old = self.win.len();
                self.win.resize(want, 0);
                let n = self.src.read(&mut self.win[old..])?;
                self.win.truncate(old + n);
                if n == 0 { self.eof = true; }
            }
            let ctx = self.ctx.lock().unwrap();
            let mut out = ptr::null_mut();
            let mut consumed = 0u64;
            ctx.check(unsafe { sys::ntg_parse_fastx_chunk(ctx.raw, self.win.as_ptr(), self.win.len(), self.format, self.eof as i32, &mut out, &mut consumed) })?;
            let r = unsafe { &*out };
            if !self.eof && consumed == 0 && r.error.kind == 0 {      // not one complete record yet: a larger window
                unsafe { sys::ntg_records_free(out) };
                want *= 2;
                continue;
            }
            self.format = r.format;
            self.recs = out; self.next_row = 0; self.consumed = consumed;
            return Ok(());
        }
    }
}
impl<R: Read + Send> FastxReader for GpuReader<R> {
    fn next(&mut self) -> Option<Result<SequenceRecord<'_>, ParseError>> {
        loop {
            if self.done { return None; }
            let r = unsafe { &*self.recs };
            if self.next_row < r.n_records {
                let row = unsafe { *r.records.add(self.next_row as usize) };
                self.next_row += 1;
                self.position = Position::new(row.line + self.line_base, row.start + self.byte_base);
                if self.line_ending.is_none() { self.line_ending = match r.line_ending { 2 => Some(LineEnding::Windows), 1 => Some(LineEnding::Unix), _ => None }; }
                let rec = SequenceRecord { buf: &self.win, row, format: format_of(r.format).unwrap(), position: self.position.clone(),
                                           line_ending: self.line_ending.unwrap_or(LineEnding::Unix) };
                return Some(Ok(rec));
            }
            if r.error.kind != 0 { self.done = true; return Some(Err(parse_error_from(&r.error, self.line_base))); }
            if self.eof { self.done = true; return None; }
            if let Err(e) = self.refill() { self.done = true; return Some(Err(e)); }
        }
    }
    fn position(&self) -> &Position { &self.position }
    fn line_ending(&self) -> Option<LineEnding> { self.line_ending }
}
impl<R: Read + Send> Drop for GpuReader<R> { fn drop(&mut self) { if !self.recs.is_null() { unsafe { sys::ntg_records_free(self.recs) } } } }

/// `parse_fastx_reader` (`src/parser/mod.rs:85-150`): two magic bytes pick a decompressor (gzip here; plug flate2 / bzip2 / xz /
/// zstd readers in exactly as the reference does), the first decompressed byte picks the format.
pub fn parse_fastx_reader<'a, R: 'a + Read + Send>(mut reader: R) -> Result<Box<dyn FastxReader + 'a>, ParseError> {
    let mut magic = [0u8; 2];
    reader.read_exact(&mut magic).map_err(|_| empty_file())?;
    let chained = io::Cursor::new(magic).chain(reader);
    // (the crate's own `flate2::read::MultiGzDecoder::new(chained)` etc. go here, unchanged: decompression is host work)
    let mut rd = GpuReader { ctx: default_context()?, src: chained, win: Vec::new(), eof: false, format: sys::NTG_FMT_NONE, recs: ptr::null_mut(),
                             next_row: 0, consumed: 0, byte_base: 0, line_base: 0, position: Position::new(1, 0), line_ending: None, done: false };
    rd.refill()?;
    let e = unsafe { &(*rd.recs).error };
    if unsafe { (*rd.recs).n_records } == 0 && (e.kind == sys::NTG_EEMPTY_FILE || e.kind == sys::NTG_EUNKNOWN_FORMAT) {
        return Err(parse_error_from(e, 0));                      // fails up front, like get_fastx_reader (mod.rs:37-46)
    }
    Ok(Box::new(rd))
}
pub fn parse_fastx_file<P: AsRef<Path>>(path: P) -> Result<Box<dyn FastxReader>, ParseError> { parse_fastx_reader(File::open(path)?) }
pub fn parse_fastx_stdin() -> Result<Box<dyn FastxReader>, ParseError> { parse_fastx_reader(io::stdin()) }
fn empty_file() -> ParseError { ParseError { msg: "Failed to read the first two bytes. Is the file empty?".into(), kind: ParseErrorKind::EmptyFile, position: ErrorPosition::default(), format: None } }

// ------------------------------------------------------------------------------------------------ Sequence (src/sequence.rs:156-253)
/// Items of `canonical_kmers` / `bit_kmers` for one sequence, produced by one device call.
pub struct KmerItems { pub pos: Vec<u32>, pub was_rc: Vec<bool>, pub val_lo: Vec<u64>, pub val_hi: Option<Vec<u64>> }

fn one_batch(seq: &[u8]) -> [u64; 2] { [0, seq.len() as u64] }
fn take_items(it: *mut sys::ntg_items) -> KmerItems {
    let r = unsafe { &*it };
    let n = r.n_items as usize;
    let sl = |p: *const u64| if p.is_null() { None } else { Some(unsafe { std::slice::from_raw_parts(p, n) }.to_vec()) };
    let out = KmerItems {
        pos: if n == 0 { vec![] } else { unsafe { std::slice::from_raw_parts(r.pos, n) }.to_vec() },
        was_rc: if r.was_rc.is_null() || n == 0 { vec![false; n] } else { unsafe { std::slice::from_raw_parts(r.was_rc, n) }.iter().map(|&b| b != 0).collect() },
        val_lo: sl(r.val_lo).unwrap_or_default(),
        val_hi: sl(r.val_hi),
    };
    unsafe { sys::ntg_items_free(it) };
    out
}

pub trait Sequence<'a> {
    fn sequence(&'a self) -> &'a [u8];

    /// `strip_returns` (`sequence.rs:165-191`): borrowed when nothing had to be removed
    fn strip_returns(&'a self) -> Cow<'a, [u8]> { xform(self.sequence(), None) }
    /// `normalize` (`sequence.rs:226-232`)
    fn normalize(&'a self, iupac: bool) -> Cow<'a, [u8]> { xform(self.sequence(), Some(iupac)) }
    /// `reverse_complement` (`sequence.rs:202-208`)
    fn reverse_complement(&'a self) -> Vec<u8> {
        let s = self.sequence();
        let mut out = vec![0u8; s.len()];
        let ctx = default_context().expect("libntgpu context");
        let ctx = ctx.lock().unwrap();
        ctx.check(unsafe { sys::ntg_reverse_complement(ctx.raw, s.as_ptr(), one_batch(s).as_ptr(), 1, out.as_mut_ptr()) }).expect("ntg_reverse_complement");
        out
    }
    /// `canonical_kmers(k, &rc)` (`sequence.rs:237-239`): `(pos, slice, was_rc)` with the slice taken from `self` or from `rc`
    fn canonical_kmers(&'a self, k: u8, reverse_complement: &'a [u8]) -> Box<dyn Iterator<Item = (usize, &'a [u8], bool)> + 'a> {
        let s = self.sequence();
        let ctx = default_context().expect("libntgpu context");
        let ctx = ctx.lock().unwrap();
        let mut it = ptr::null_mut();
        ctx.check(unsafe { sys::ntg_canonical_kmers(ctx.raw, s.as_ptr(), reverse_complement.as_ptr(), one_batch(s).as_ptr(), 1, k as u32, &mut it) }).expect("ntg_canonical_kmers");
        let items = take_items(it);
        let (k, len) = (k as usize, s.len());
        Box::new(items.pos.into_iter().zip(items.was_rc).map(move |(p, rc)| {
            let p = p as usize;
            if rc { (p, &reverse_complement[len - p - k..len - p], true) } else { (p, &s[p..p + k], false) }      // kmer.rs:121-128
        }))
    }
    /// `kmers(k)` (`sequence.rs:245-247`): every window — positions come from `ntg_kmers`, the items are slices of `self`
    fn kmers(&'a self, k: u8) -> Box<dyn Iterator<Item = &'a [u8]> + 'a> {
        let s = self.sequence();
        let ctx = default_context().expect("libntgpu context");
        let ctx = ctx.lock().unwrap();
        let mut it = ptr::null_mut();
        ctx.check(unsafe { sys::ntg_kmers(ctx.raw, s.as_ptr(), one_batch(s).as_ptr(), 1, k as u32, &mut it) }).expect("ntg_kmers");
        let items = take_items(it);
        Box::new(items.pos.into_iter().map(move |p| &s[p as usize..p as usize + k as usize]))
    }
    /// `bit_kmers(k, canonical)` (`sequence.rs:250-252`): `(pos, (kmer, k), was_rc)`
    fn bit_kmers(&'a self, k: u8, canonical: bool) -> Box<dyn Iterator<Item = (usize, (u64, u8), bool)> + 'a> {
        let s = self.sequence();
        let ctx = default_context().expect("libntgpu context");
        let ctx = ctx.lock().unwrap();
        let mut it = ptr::null_mut();
        ctx.check(unsafe { sys::ntg_bit_kmers(ctx.raw, s.as_ptr(), one_batch(s).as_ptr(), 1, k as u32, canonical as i32, &mut it) }).expect("ntg_bit_kmers");
        let items = take_items(it);
        Box::new(items.pos.into_iter().zip(items.val_lo).zip(items.was_rc).map(move |((p, v), rc)| (p as usize, (v, k), rc)))
    }
}
impl<'a> Sequence<'a> for &'a [u8] { fn sequence(&'a self) -> &'a [u8] { self } }
impl<'a> Sequence<'a> for [u8] { fn sequence(&'a self) -> &'a [u8] { self } }
impl<'a> Sequence<'a> for Cow<'a, [u8]> { fn sequence(&'a self) -> &'a [u8] { self } }

fn xform(s: &[u8], iupac: Option<bool>) -> Cow<'_, [u8]> {
    let ctx = default_context().expect("libntgpu context");
    let ctx = ctx.lock().unwrap();
    let mut out = vec![0u8; s.len().max(1)];
    let (mut out_offs, mut changed) = ([0u64; 2], [0u8; 1]);
    let offs = one_batch(s);
    let st = unsafe {
        match iupac {
            Some(f) => sys::ntg_normalize(ctx.raw, s.as_ptr(), offs.as_ptr(), 1, f as i32, out.as_mut_ptr(), out_offs.as_mut_ptr(), changed.as_mut_ptr()),
            None => sys::ntg_strip_returns(ctx.raw, s.as_ptr(), offs.as_ptr(), 1, out.as_mut_ptr(), out_offs.as_mut_ptr(), changed.as_mut_ptr()),
        }
    };
    ctx.check(st).expect("ntg_normalize / ntg_strip_returns");
    if changed[0] == 0 { Cow::Borrowed(s) } else { out.truncate(out_offs[1] as usize); Cow::Owned(out) }
}

/// `bitkmer::{reverse_complement, canonical, minimizer}` (`src/bitkmer.rs:112-162`) on slices of k-mers.
pub mod bitkmer {
    use super::*;
    pub type BitKmerSeq = u64;
    pub type BitKmer = (BitKmerSeq, u8);
    pub fn minimizers(kmers: &[u64], k: u8, m: u8) -> Vec<u64> {
        let mut out = vec![0u64; kmers.len()];
        let ctx = default_context().expect("libntgpu context");
        let ctx = ctx.lock().unwrap();
        ctx.check(unsafe { sys::ntg_bitkmer_minimizer(ctx.raw, kmers.as_ptr(), kmers.len(), k as u32, m as u32, out.as_mut_ptr()) }).expect("ntg_bitkmer_minimizer");
        out
    }
    pub fn canonical(kmers: &[u64], k: u8) -> (Vec<u64>, Vec<bool>) {
        let (mut out, mut rc) = (vec![0u64; kmers.len()], vec![0u8; kmers.len()]);
        let ctx = default_context().expect("libntgpu context");
        let ctx = ctx.lock().unwrap();
        ctx.check(unsafe { sys::ntg_bitkmer_canonical(ctx.raw, kmers.as_ptr(), kmers.len(), k as u32, out.as_mut_ptr(), rc.as_mut_ptr()) }).expect("ntg_bitkmer_canonical");
        (out, rc.into_iter().map(|b| b != 0).collect())
    }
}

// ------------------------------------------------------------------------------------------------ the fused pass
/// What the README loop computes (`src/lib.rs:15-36`), in one pass over a file (plain or gzip / BGZF) — `ntg_tally_fastx_file`.
pub fn tally_fastx_file<P: AsRef<Path>>(path: P, k: u8, m: u8, inflate_threads: i32) -> Result<(sys::ntg_tallies, Option<ParseError>), ParseError> {
    let ctx = default_context()?;
    let ctx = ctx.lock().unwrap();
    let cfg = sys::ntg_tally_config { k: k as u32, m: m as u32, allow_iupac: 0, has_query: 0, query: [0; 64], flags: 0, qmask_score: 0 };
    let (mut t, mut e) = (sys::ntg_tallies::default(), unsafe { std::mem::zeroed::<sys::ntg_parse_error>() });
    let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|_| empty_file())?;
    ctx.check(unsafe { sys::ntg_tally_fastx_file(ctx.raw, c.as_ptr(), &cfg, inflate_threads, &mut t, &mut e) })?;
    Ok((t, if e.kind != 0 { Some(parse_error_from(&e, 0)) } else { None }))
}
/// The same over any `Read` (stdin, sockets): pieces are staged in pinned memory by the library (`ntg_stream_*`).
pub fn tally_fastx_reader<R: Read>(mut reader: R, k: u8, m: u8) -> Result<(sys::ntg_tallies, Option<ParseError>), ParseError> {
    let ctx = default_context()?;
    let ctx = ctx.lock().unwrap();
    let cfg = sys::ntg_tally_config { k: k as u32, m: m as u32, allow_iupac: 0, has_query: 0, query: [0; 64], flags: 0, qmask_score: 0 };
    let mut s = ptr::null_mut();
    ctx.check(unsafe { sys::ntg_stream_open(ctx.raw, &cfg, &mut s) })?;
    let res = (|| {
        loop {
            let (mut p, mut avail) = (ptr::null_mut(), 0usize);
            ctx.check(unsafe { sys::ntg_stream_acquire(s, &mut p, &mut avail) })?;
            let n = reader.read(unsafe { std::slice::from_raw_parts_mut(p, avail) })?;      // straight into the pinned staging buffer
            if n == 0 { break; }
            ctx.check(unsafe { sys::ntg_stream_commit(s, n) })?;
        }
        let (mut t, mut e) = (sys::ntg_tallies::default(), unsafe { std::mem::zeroed::<sys::ntg_parse_error>() });
        ctx.check(unsafe { sys::ntg_stream_finish(s, &mut t, &mut e) })?;
        Ok((t, if e.kind != 0 { Some(parse_error_from(&e, 0)) } else { None }))
    })();
    unsafe { sys::ntg_stream_close(s) };
    res
}
