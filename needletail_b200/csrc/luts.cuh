// luts.cuh — byte-class tables (host construction + upload).  Part of the unity build.
#pragma once
#include "common.cuh"

struct HostLuts {
    uint8_t norm[2][256], comp[256], code[256], ncls[256];
    HostLuts() {
        // sequence::normalize, src/sequence.rs:19-62
        for (int iu = 0; iu < 2; iu++)
            for (int b = 0; b < 256; b++) {
                uint8_t o;
                switch (b) {
                    case 'A': case 'C': case 'G': case 'T': case 'N': case '-': o = (uint8_t)b; break;
                    case 'a': o = 'A'; break; case 'c': o = 'C'; break; case 'g': o = 'G'; break;
                    case 't': case 'u': case 'U': o = 'T'; break;
                    case '.': case '~': o = '-'; break;
                    case 'B': case 'D': case 'H': case 'V': case 'R': case 'Y': case 'S': case 'W': case 'K': case 'M':
                        o = iu ? (uint8_t)b : (uint8_t)'N'; break;
                    case 'b': case 'd': case 'h': case 'v': case 'r': case 'y': case 's': case 'w': case 'k': case 'm':
                        o = iu ? (uint8_t)(b - 32) : (uint8_t)'N'; break;
                    case ' ': case '\t': case '\r': case '\n': o = 0; break;   // deleted
                    default: o = 'N'; break;
                }
                norm[iu][b] = o;
            }
        // sequence::complement, src/sequence.rs:67-105
        for (int b = 0; b < 256; b++) comp[b] = (uint8_t)b;
        const char* from = "acgtACGTrykmbvdhswRYKMBVDHSW";
        const char* to   = "tgcaTGCAyrmkvbhdswYRMKVBHDSW";
        for (int i = 0; from[i]; i++) comp[(uint8_t)from[i]] = (uint8_t)to[i];
        // bitkmer::NUC2BIT_LOOKUP (src/bitkmer.rs:5-18) == kmer::is_good_base set (src/kmer.rs:6-8)
        for (int b = 0; b < 256; b++) code[b] = 4;
        code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2; code['T'] = code['t'] = 3;
        // class after normalize(): the iupac flag never turns a byte into ACGT, so one table serves both
        for (int b = 0; b < 256; b++) {
            uint8_t o = norm[0][b];
            ncls[b] = (o == 0) ? 5 : (o == 'A' ? 0 : o == 'C' ? 1 : o == 'G' ? 2 : o == 'T' ? 3 : 4);
        }
    }
};
static const HostLuts& host_luts() { static HostLuts l; return l; }

static int ntg_upload_luts(ntg_ctx* ctx) {
    const HostLuts& l = host_luts();
    NTG_CUDA(ctx, cudaMemcpyToSymbol(c_norm, l.norm, sizeof(l.norm)));
    NTG_CUDA(ctx, cudaMemcpyToSymbol(c_comp, l.comp, sizeof(l.comp)));
    NTG_CUDA(ctx, cudaMemcpyToSymbol(c_code, l.code, sizeof(l.code)));
    NTG_CUDA(ctx, cudaMemcpyToSymbol(c_ncls, l.ncls, sizeof(l.ncls)));
    return NTG_OK;
}
