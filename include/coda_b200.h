/* coda_b200 -- C ABI of the B200-native CODA acquisition hot path.
 *
 * Drop-in boundary.  The reference (justinkay/coda) is pure Python/PyTorch
 * and has no FFI of its own; these entry points are what a ctypes binding inside
 * coda/coda.py would call in place of the ATen op chains on the acquisition path.  Each
 * declaration cites the reference lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises;
 *   - return value: CODA_B200_OK or a negative error code, message via coda_b200_last_error();
 *   - numerical problems are reported through a device-side `flags` word (bits below) that the
 *     caller reads at its next host sync -- the reference raises RuntimeError('[NUMERIC ERROR]')
 *     from coda/util.py:17-25 at the same places;
 *   - layouts: preds [H][N][C] fp32 (coda/datasets.py:14); D (dirichlets) [H][C][C] fp32;
 *     U (un-normalised pi_hat_xi) [N][C] fp32; hard [N][H] u16; Hp = H rounded up to 32.
 *   - built for sm_100a only.
 */
#ifndef CODA_B200_H
#define CODA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CODA_B200_VERSION 100
#define CODA_B200_NODES 256 /* quadrature nodes, coda/coda.py:79 */

#define CODA_B200_OK 0
#define CODA_B200_EINVAL (-1)
#define CODA_B200_ECUDA (-2)

/* device-side flag bits */
#define CODA_B200_FLAG_NONFINITE_INPUT 0x01u /* NaN/Inf in preds */
#define CODA_B200_FLAG_RANGE_INPUT 0x02u     /* preds outside [0, 1]: not post-softmax scores */
#define CODA_B200_FLAG_NONFINITE_TABLE 0x04u /* util._check(pdf/cdf/integrand), coda.py:96-112 */
#define CODA_B200_FLAG_NONFINITE_PI 0x08u    /* pi_hat_xi row sum not finite */
#define CODA_B200_FLAG_NONFINITE_PBEST 0x10u /* util._check(pbest), coda.py:330 */
#define CODA_B200_FLAG_NONFINITE_EIG 0x20u   /* util._check(Pbest(beta) normalized), coda.py:115 */

typedef void* coda_stream_t;

/* ---- plumbing ---------------------------------------------------------------------- */
const char* coda_b200_last_error(void);
int coda_b200_version(void);
int coda_b200_sm_count(void);
int coda_b200_device_check(void); /* fails loudly when no sm_100 device is present */

/* ---- construction (coda/coda.py:172-203) --------------------------------------------- */

/* One pass over the slab: per-model argmax (coda.py:217, 263, 316), ensemble-mean pseudo
 * label (coda/util.py:13-14 + coda.py:193-194), unanimity bit (coda.py:215-219). */
/* ens_out (optional) [N][C]: E[n][c] = sum_h preds[h][n][c], the un-normalised ensemble of coda/util.py:13-14. */
int coda_b200_scan_slab(const float* preds, int H, int64_t N, int C, uint16_t* hard, int32_t* pseudo,
                        uint8_t* disagree, float* ens_out, uint32_t* flags, coda_stream_t stream);

/* Soft confusion sums, coda.py:42 einsum('nc,hnj->hcj').  conf_fx [H][C][C] int64 fixed point
 * (value * 2^fx_shift), ACCUMULATED into; exact and order-independent so shards can be summed. */
int coda_b200_confusion_accum(const float* preds, const int32_t* pseudo, int H, int64_t N, int C, int fx_shift,
                              int64_t* conf_fx, coda_stream_t stream);

/* Same sums with the items visited in pseudo-label order (`order` = any permutation that groups equal
 * pseudo labels; C <= 128): register accumulation, no shared-memory atomics.  Bit-identical result. */
int coda_b200_confusion_sorted(const float* preds, const int32_t* pseudo, const int32_t* order, int H, int64_t N,
                               int C, int fx_shift, int64_t* conf_fx, coda_stream_t stream);

/* Row-normalise (coda.py:43) and build the Dirichlet prior (coda.py:46-63, 196). */
int coda_b200_init_dirichlets(const int64_t* conf_fx, int H, int C, int fx_shift, double prior_strength,
                              double multiplier, int uniform_prior, float* D, coda_stream_t stream);

/* ---- consensus marginals (CODA.update_pi_hat, coda.py:226-233) ------------------------ */

/* U[n][c] = sum_h sum_s D[h][c][s] preds[h][n][s]  (coda.py:227-229, `adjusted` never stored). */
int coda_b200_pi_full(const float* preds, const float* D, int H, int64_t N, int C, float* U, coda_stream_t stream);

/* Row-normalise with the 1e-12 clamp (coda.py:230) and accumulate sum_n pi_hat_xi[n][:]
 * (coda.py:232) into pisum_fx [C] (int64 fixed point, ACCUMULATED).  xi_out may be NULL. */
int coda_b200_pi_reduce(float* U, int64_t N, int C, int fx_shift, float* xi_out, int64_t* pisum_fx, uint32_t* flags,
                        coda_stream_t stream);

/* ---- posterior update (CODA.add_label, coda.py:315-319) ------------------------------- */
/* sel = {local item index or -1 when another shard owns it, revealed class}. */

/* jvec[h] = p_h(idx) (coda.py:316) and labeled[idx] = 1 (coda.py:323) on the owner shard; jvec = 0 elsewhere,
 * so a SUM all-reduce of jvec hands the owner's row to every shard. */
int coda_b200_label_row(const uint16_t* hard, int H, int64_t N, const int64_t* sel, int32_t* jvec, uint8_t* labeled,
                        coda_stream_t stream);
/* D[h][t][jvec[h]] += lr  (coda.py:317). */
int coda_b200_label_apply(float* D, int H, int C, const int64_t* sel, const int32_t* jvec, double lr,
                          coda_stream_t stream);
/* Optional class-major shadow copy T[s][c][n] = preds[model_of_slot[s]][n][c] for S of the H models (no
 * reference counterpart: a layout for the one-float-per-(model, item) gather of the rank-1 refresh). */
int coda_b200_shadow_build(const float* preds, int H, int64_t N, int C, const int32_t* model_of_slot, int S, float* T,
                           coda_stream_t stream);
/* update_pi_hat after the rank-1 change of D (coda.py:319): U[n][t] += lr * sum_h preds[h][n][jvec[h]],
 * then the same normalise + column sums as pi_reduce.  With ens != NULL (scan_slab's ens_out) the sum over
 * models is taken as E[n][t'] + corrections for the models that disagree with the majority class t' of
 * jvec (exact algebra, fewer gathers).  shadow/slot_of_model (optional): models with slot_of_model[h] >= 0
 * are read from the shadow copy.  terms: int32 scratch [2 + 8*H], 8-byte aligned.  pisum_fx is ZEROED and then
 * accumulated into (one shard's sums).  ctas_per_sm (1..8)
 * bounds the grid so a concurrent stream keeps SM resources. */
int coda_b200_pi_rank1(const float* preds, const float* ens, const float* shadow, const int32_t* slot_of_model, int H,
                       int64_t N, int C, const int64_t* sel, const int32_t* jvec, double lr, int fx_shift,
                       int32_t* terms, float* U, int64_t* pisum_fx, uint32_t* flags, int ctas_per_sm,
                       coda_stream_t stream);
/* cudaLimitMaxL2FetchGranularity hint (32/64/128 B) for the sector-gather kernels. */
int coda_b200_set_l2_fetch_granularity(int bytes);

/* ---- Beta quadrature tables (dirichlet_to_beta coda.py:14-25, compute_pbest_beta_batched
 *      coda.py:77-119, batch_update_beta coda.py:150-168) for classes [cls_lo, cls_hi) ------- */
size_t coda_b200_tables_scratch_bytes(int H, int ncls);
/* sel (optional, device): {idx, class}; when non-NULL exactly one class, sel[1], is rebuilt (host-free loop). */
/* dLb / Gb (optional, both or neither): the same tables as bf16 limbs in tensor-core operand order for
 * coda_b200_pair_rows_tc: dLb [C][Hp/32][3][256*32], Gb [C][16][4][Hp*16] bf16, zero-initialised by the caller. */
int coda_b200_beta_tables(const float* D, const float* grid_x, int H, int C, int P, double hyp_w, int cls_lo,
                          int cls_hi, const int64_t* sel, void* scratch, float* dL /*[C][H][P]*/,
                          float* G0T /*[C][P][Hp]*/, float* G1T /*[C][P][Hp]*/, float* PB /*[C][Hp]*/, void* dLb,
                          void* Gb, uint32_t* flags, coda_stream_t stream);

/* pi_hat (coda.py:232-233), P(best) vector m0 == get_pbest() (coda.py:253, 325-332), H_before
 * (coda.py:254) and argmax (coda.py:346). */
int coda_b200_mixture(const int64_t* pisum_fx, const float* PB, int H, int C, float* pi_hat, float* m0 /*[Hp]*/,
                      float* h_before, int64_t* best_model, uint32_t* flags, coda_stream_t stream);

/* ---- hypothetical-update pairs (eig_batched inner loop, coda.py:261-279) ---------------- */
int coda_b200_pair_count(const uint16_t* hard, int H, int64_t N, int C, int32_t* ent_cnt /*[N]*/,
                         int32_t* cls_heavy /*[C], accumulated*/, coda_stream_t stream);
int coda_b200_pair_fill(const uint16_t* hard, int H, int64_t N, int C, const int64_t* ent_off /*[N+1]*/,
                        const int64_t* cls_base /*[C+1]*/, int32_t* cls_cursor /*[C] zeroed*/, int32_t* ent_pair,
                        uint16_t* ent_cls, uint32_t* zmask /*[npairs][Hp/32]*/, uint16_t* pair_cls,
                        int32_t* pair_item, coda_stream_t stream);
/* tiles [ntiles][4] int32 = {class, first pair id, count <= 32, 0}; processes tiles [tile_lo, tile_hi).
 * Writes gain[pair] = H_before - H_after (coda.py:274-276) and, if ph_cache != NULL, the normalised
 * P(best | hypothetical) row (coda.py:271-273).  With sel != NULL the launch covers [0, tile_hi - tile_lo)
 * tiles of class sel[1] (pass the largest per-class tile count). */
int coda_b200_pair_rows(const int32_t* tiles, int tile_lo, int tile_hi, const uint32_t* zmask, const float* dL,
                        const float* G0T, const float* G1T, const float* PB, const float* m0, const float* pi_hat,
                        int H, float* ph_cache, float* gain, const int64_t* sel /*optional*/,
                        const int64_t* tile_off /*[C+1], with sel*/, uint32_t* flags, coda_stream_t stream);
/* The same computation on the tcgen05 tensor cores (Hp <= 256): tiles128 are tiles of <= 128 same-class pairs,
 * operands come from the bf16 limb tables of coda_b200_beta_tables. */
int coda_b200_pair_rows_tc(const int32_t* tiles128, int tile_lo, int tile_hi, const uint32_t* zmask, const void* dLb,
                           const void* Gb, const float* PB, const float* m0, const float* pi_hat, int H,
                           float* ph_cache, float* gain, const int64_t* sel, const int64_t* tile_off, uint32_t* flags,
                           coda_stream_t stream);
/* gain for every pair from cached rows (coda.py:274-276 only).  filter: 0 = all pairs, 1 = all but class t,
 * 2 = only class t, with t = sel[1] when sel != NULL (device) else cls_host; cls_base [C+1] gives the pair-id
 * range of every class.  (1 then 2 lets the class-t row refresh overlap the rest of the stream.) */
int coda_b200_pair_gain(const float* ph_cache, const uint16_t* pair_cls, int64_t npairs, int H, const float* PB,
                        const float* m0, const float* pi_hat, float* gain, const int64_t* sel,
                        const int64_t* cls_base, int cls_host, int filter, coda_stream_t stream);

/* ---- selection (coda.py:278, get_next_item_to_label coda.py:283-313) -------------------- */
int coda_b200_eig_blocks(int64_t N); /* number of partial records eig_points writes */
/* ell (optional): [N][ell_k] x {pair id, class} copy of the CSR lists from coda_b200_ell_build, ell_k >= the
 * longest list and <= 32; removes one dependent load level. */
int coda_b200_eig_points(const float* U, int64_t N, int C, const int64_t* ent_off, const int32_t* ent_pair,
                         const uint16_t* ent_cls, const float* gain, const int64_t* cls_base, const uint8_t* labeled,
                         const uint8_t* disagree, int64_t n_offset, const int32_t* ell, int ell_k, float* eig,
                         int64_t* partials /*[blocks][5]*/, uint32_t* flags, coda_stream_t stream);
int coda_b200_ell_build(const int64_t* ent_off, const int32_t* ent_pair, const uint16_t* ent_cls, int64_t N, int K,
                        int32_t* ell /*[N][K][2]*/, coda_stream_t stream);
int coda_b200_select_merge(const int64_t* recs, int nrec, int64_t* out /*[5]*/, coda_stream_t stream);
int coda_b200_ties(const float* eig, int64_t N, const uint8_t* labeled, const uint8_t* disagree, int64_t n_offset,
                   const int64_t* best /*[5]*/, int cap, int64_t* tie_hdr /*[2]*/, int64_t* tie_idx, float* tie_val,
                   coda_stream_t stream);
/* device-resident oracle stand-in (coda/oracle.py:23-24) for host-free benchmark loops. */
int coda_b200_device_pick(const int64_t* tie_hdr /*or NULL*/, const int64_t* best /*merged record*/,
                          const int64_t* labels_global, int64_t n_offset, int64_t N, const float* eig, int64_t* sel,
                          int64_t* hist_idx, float* hist_q, int64_t step, coda_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CODA_B200_H */
