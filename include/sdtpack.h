/*
 * sdtpack.h — multi-threaded FASTA/FASTQ record parser + 2-bit packer (host only, no CUDA).
 *
 * once the hashing kernels run at GPU speed the reference's byte-wise,
 * single-threaded parser (readseqInBuf readseq1by1.c:122-178, readseqfq :281-340, driven by the
 * AIO loop prlHashReads.c:432-613) is the wall-clock bottleneck of the stage.  This reader produces
 * exactly what that code leaves in seqBuffer/lenBuffer, already packed for sdtgpu_push_reads:
 *   - records in the reference's arrival order: paired files alternate read1, read2
 *     (prlHashReads.c:493-567); when one file of a pair ends first the other is drained;
 *   - per record only the FIRST sequence line is taken; the first max_read_len characters of it are
 *     considered; N/n -> 4 with n_kmer (else via base2int -> 3 = G); a-z folded to upper case;
 *     letters -> base2int ((c & 6) >> 1: A0 C1 T2 G3, inc/def.h:39); '.' -> A; anything else
 *     dropped (readseq1by1.c:147-171);
 *   - reverse != 0: the read is reverse-complemented (reverse2k, readseq1by1.c:749-764).
 * Unlike the reference's loop, reads shorter than K+1 are NOT skipped here: they are emitted with
 * their length and the device kernel skips them (same table; ordinals stay order-preserving).
 * Output layout = include/sdtgpu.h: tight strings (seq.c:49-90), stride_bytes per read, lens[],
 * optional N mask (stride_bytes / 2 bytes per read, bit 7 of byte 0 = base 0).
 */
#ifndef SDTPACK_H
#define SDTPACK_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdtpack_reader sdtpack_reader;

/* path2 == NULL: single-ended stream.  fastq: 0 = FASTA ('>'), 1 = FASTQ ('@', 4 lines per record).
 * n_threads <= 0: one per online core (at most 64).  Returns 0, or -1 (errno set) if a file cannot
 * be opened/mapped. */
int sdtpack_open (sdtpack_reader **out, const char *path1, const char *path2, int fastq, int n_threads);

/* Parses up to max_reads records (an even number is consumed for pairs) into the caller's buffers.
 * nmask may be NULL when n_kmer == 0.  Returns the number of records written, 0 at end of input. */
int64_t sdtpack_next (sdtpack_reader *r, int max_read_len, int n_kmer, int reverse,
		      uint8_t *packed, uint32_t *lens, uint8_t *nmask, uint64_t max_reads, uint32_t stride_bytes);

void sdtpack_close (sdtpack_reader *r);

#ifdef __cplusplus
}
#endif
#endif
