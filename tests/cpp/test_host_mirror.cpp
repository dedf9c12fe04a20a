// The reference's own tests, re-stated against the C++ host mirror (needletail_b200/host/needletail.hpp),
// i.e. through the C ABI on the GPU.  Each block names the Rust test it mirrors.
//   g++ -std=c++17 -O1 tests/cpp/test_host_mirror.cpp -o /tmp/test_host_mirror -L needletail_b200 -lntgpu
#include <cassert>
#include <cstdio>
#include <cstring>

#include "../../needletail_b200/host/needletail.hpp"

using namespace needletail;
static Bytes B(const char* s) { return Bytes(s, s + std::strlen(s)); }
static ByteView V(const Bytes& b) { return ByteView(b.data(), b.size()); }
static bool eq(ByteView v, const char* s) { return v.size() == std::strlen(s) && std::memcmp(v.data(), s, v.size()) == 0; }
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    // src/parser/fastq.rs:473-511  test_simple_fastq (both line endings)
    for (int crlf = 0; crlf < 2; crlf++) {
        auto reader = parse_fastx_reader(B(crlf ? "@test\r\nAGCT\r\n+test\r\n~~a!\r\n@test2\r\nTGCA\r\n+test\r\nWUI9"
                                                : "@test\nAGCT\n+test\n~~a!\n@test2\nTGCA\n+test\nWUI9"));
        int i = 0;
        while (auto rec = reader->next()) {
            auto& r = std::get<SequenceRecord>(*rec);
            if (i == 0) { CHECK(eq(r.id(), "test")); CHECK(eq(r.raw_seq(), "AGCT")); CHECK(eq(*r.qual(), "~~a!")); }
            else { CHECK(eq(r.id(), "test2")); CHECK(eq(r.raw_seq(), "TGCA")); CHECK(eq(*r.qual(), "WUI9")); }
            CHECK(*reader->line_ending() == (crlf ? LineEnding::Windows : LineEnding::Unix));
            i++;
        }
        CHECK(i == 2);
    }
    {   // src/parser/fastq.rs:514-532  EOF in qual => UnequalLengths ; EOF in seq => UnexpectedEnd
        auto r1 = parse_fastx_reader(B("@test\nACGT\n+\nIII"));
        auto e = r1->next(); CHECK(e && std::get<ParseError>(*e).kind == ParseErrorKind::UnequalLengths);
        auto r2 = parse_fastx_reader(B("@test\nAGCT\n+test\n~~a!\n@test2\nTGCA"));
        CHECK(std::holds_alternative<SequenceRecord>(*r2->next()));
        auto e2 = r2->next(); CHECK(e2 && std::get<ParseError>(*e2).kind == ParseErrorKind::UnexpectedEnd);
        CHECK(!r2->next());
    }
    {   // src/parser/fasta.rs:409-425  wrapped FASTA: raw_seq keeps '\n', num_bases drops it
        auto reader = parse_fastx_reader(B(">test\nACGT\nACGT\n>test2\nTGCA\nTG"));
        auto a = std::get<SequenceRecord>(*reader->next());
        CHECK(eq(a.id(), "test")); CHECK(eq(a.raw_seq(), "ACGT\nACGT")); CHECK(a.num_bases() == 8);
        auto b = std::get<SequenceRecord>(*reader->next());
        CHECK(eq(b.raw_seq(), "TGCA\nTG")); CHECK(b.num_bases() == 6); CHECK(!reader->next());
    }
    {   // src/parser/mod.rs:182-200  < 2 bytes => EmptyFile ; unknown first byte => UnknownFormat
        try { parse_fastx_reader(B("@")); CHECK(false); } catch (const ParseError& e) { CHECK(e.kind == ParseErrorKind::EmptyFile); }
        try { parse_fastx_reader(B("Not a valid file")); CHECK(false); } catch (const ParseError& e) { CHECK(e.kind == ParseErrorKind::UnknownFormat); }
    }
    {   // src/sequence.rs:197-201,215-225,316-344  doc-tests of the Sequence trait
        Bytes s = B("AACC"); CHECK(eq(V(Sequence(V(s)).reverse_complement()), "GGTT"));
        Bytes t = B("ADGH"); CHECK(eq(V(Sequence(V(t)).normalize(false)), "ANGN")); CHECK(eq(V(Sequence(V(t)).normalize(true)), "ADGH"));
        Bytes u = B("ACGU"); CHECK(eq(V(Sequence(V(u)).normalize(true)), "ACGT"));
        Bytes w = B("ACGT\r\nAC"); CHECK(eq(V(Sequence(V(w)).strip_returns()), "ACGTAC"));
    }
    {   // src/kmer.rs:210-225  canonical_kmers skips the N and reports positions 0 and 3
        Bytes s = B("AGNTA"); Bytes rc = Sequence(V(s)).reverse_complement();
        auto it = Sequence(V(s)).canonical_kmers(2, V(rc));
        CHECK(it.size() == 2 && it[0].pos == 0 && it[1].pos == 3);
        // src/bitkmer.rs:241-246,255-266
        Bytes a = B("ACGTA"); auto bk = Sequence(V(a)).bit_kmers(3, false);
        CHECK(bk.size() == 3 && bk[0].lo == 6 && bk[1].lo == 27 && bk[2].lo == 44 && !bk[0].was_rc);
        CHECK(bitkmer::reverse_complement(0b00011011, 4) == 0b00011011);
        CHECK(bitkmer::minimizer(0b001011, 3, 2) == 0b0010);
        CHECK(bitkmer::canonical(0b00011011, 4) == std::make_pair((uint64_t)0b00011011, false));
    }
    if (argc > 1) {
        // README / src/lib.rs:15-36 loop on tests/data/28S.fasta (C1): 738 580 bases, 8 108 AAAAs
        auto reader = parse_fastx_file(argv[1]);
        size_t n_bases = 0, n_valid_kmers = 0, n_rec = 0;
        Batch norm_in;
        while (auto record = reader->next()) {
            auto& seqrec = std::get<SequenceRecord>(*record);
            n_bases += seqrec.num_bases(); n_rec++;
            norm_in.push(seqrec.raw_seq());
        }
        auto normed = normalize_batch(norm_in, false);               // batch form of seqrec.normalize(false)
        Batch nb; for (auto& p : normed) nb.push(V(p.first));
        for (auto& per_seq : canonical_kmers_batch(nb, 4))           // norm_seq.canonical_kmers(4, &rc)
            for (auto& km : per_seq) if (km.lo == 0) n_valid_kmers++;   // "AAAA" packs to 0
        std::printf("There are %zu bases in your file.\nThere are %zu AAAAs in your file.\n", n_bases, n_valid_kmers);
        CHECK(n_rec == 570 && n_bases == 738580 && n_valid_kmers == 8108);
    }
    std::puts("host mirror ok");
    return 0;
}
