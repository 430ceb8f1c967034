/*
 * tm_gtm.h -- C ABI of libtm_gtm.so: HOST-side verification I/O for the GTM stream.
 *
 * Not part of the GPU product path and not something the FreePascal host binds: in the reference the bitstream writer
 * (TTilingEncoder.SaveStream, tilingencoder.pas:5177-5482) stays in the host and the decoder is gtm.player.js.  This library
 * exists so that an encode can be verified end to end here (no fpc, no node, and liblzma rejects the stream's lc = 8):
 *   - LZMA "alone" codec with end marker, the chunk format of LZCompress (extern.pas:420-440),
 *   - the tilemap-item serialiser of SaveStream (DoTMI :5208-5268, SkipBlock runs :5394-5436, FrameEnd :5441-5443),
 *   - a decoder with the semantics of LoadStream (:4880-5175) / gtm.player.js:365-546.
 */
#ifndef TM_GTM_H
#define TM_GTM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* src -> one LZMA "alone" stream (props byte, dictionary size, 8 x 0xFF, range-coded data, end marker).
   Returns the size, -1 if out_cap is too small, -2 on bad parameters. */
int64_t tmh_lzma_encode(const uint8_t *src, int64_t n, int lc, int lp, int pb, uint32_t dict_size, uint8_t *out, int64_t out_cap);
/* same with an explicit number of parser threads (0 = default: up to 8).  The input is parsed in fixed 256 KB blocks by the
   threads and range-coded by the caller as blocks complete; the stream does not depend on the thread count. */
int64_t tmh_lzma_encode_mt(const uint8_t *src, int64_t n, int lc, int lp, int pb, uint32_t dict_size, uint8_t *out, int64_t out_cap,
                           int n_threads);
/* decodes ONE stream starting at src; *consumed = input bytes used (streams of a GTM file follow each other back to back).
   Returns the decoded size, -1 if out_cap is too small, -2 on corrupt input. */
int64_t tmh_lzma_decode(const uint8_t *src, int64_t n, uint8_t *out, int64_t out_cap, int64_t *consumed);

/* tilemap items of n_frames consecutive frames of one keyframe sequence -> command bytes; returns the byte count */
int64_t tmh_gtm_write_frames(const int32_t *tile_idx, const int32_t *pal_idx, const int32_t *pred_x, const int32_t *pred_y,
                             const uint8_t *is_pred, const uint8_t *mirror, int n_frames, int tiles_per_frame, const uint8_t *tiles,
                             const int32_t *use_count, int64_t n_tiles, int emit_skip_blocks, int last_is_kf_end, uint8_t *out,
                             int64_t cap);

typedef struct tmh_gtm_decoder tmh_gtm_decoder;
tmh_gtm_decoder *tmh_gtm_decoder_create(void);
void tmh_gtm_decoder_destroy(tmh_gtm_decoder *d);
int tmh_gtm_decoder_dims(tmh_gtm_decoder *d, int *w, int *h, int64_t *frames);
/* plays n raw command bytes; completed frames (packed 0x00BBGGRR, w*8 x h*8) are appended to frames_out.
   Returns the number of frames produced by this call, -1 on a malformed stream. */
int64_t tmh_gtm_decode(tmh_gtm_decoder *d, const uint8_t *s, int64_t n, int32_t *frames_out, int64_t max_frames);

/* TTilingEncoder.OptimizePalettes (tilingencoder.pas:4246-4432) over the reference's Powell minimiser (powell.pas): reorders
   the colours INSIDE each palette (palettes [n_pal][pal_size] int32 0x00BBGGRR, in place) so that the per-column standard
   deviation accumulated over all palettes is maximal; host code in the reference too (not behind a DLL).  n_threads <= 0: one
   per hardware thread.  Returns the number of outer passes, -1 on bad arguments. */
int tmh_optimize_palettes(int32_t *palettes, int n_pal, int pal_size, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
