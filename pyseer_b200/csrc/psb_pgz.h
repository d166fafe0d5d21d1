// psb_pgz.h -- parallel inflate of ONE plain gzip stream (host code, no CUDA).
//
// pyseer's variant files are usually `gzip`ed text (input.open_variant_file, pyseer/input.py:268-298):
// one serial deflate stream, which zlib inflates at ~0.4 GB/s of text however many threads parse
// behind it.  This reader inflates such a file on several threads: see psb_pgz.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct psb_pgz;

// nullptr when the file cannot be opened / mapped or does not start with a gzip header.
// chunk_bytes: compressed bytes per work item (0: default).
psb_pgz *psb_pgz_open(const char *path, int n_threads, size_t chunk_bytes);
// Next bytes of the decompressed stream; 0 at its end, -1 on a corrupt stream (psb_pgz_error).
int64_t psb_pgz_read(psb_pgz *z, char *dst, int64_t want);
const char *psb_pgz_error(const psb_pgz *z);
void psb_pgz_set_threads(psb_pgz *z, int n_threads);
// diagnostics: work items decoded / of them thrown away (false block starts, overrun by a neighbour)
void psb_pgz_stats(const psb_pgz *z, int64_t out[2]);
void psb_pgz_close(psb_pgz *z);
// One raw deflate stream of known decompressed size (a BGZF block) -> bytes, with the decoder of this
// file and without markers (no window before it); thread safe; 0 = ok, -1 = not a valid stream of that size.
int psb_pgz_inflate_exact(const unsigned char *in, size_t in_len, unsigned char *out, size_t out_len);
