/* jamun_b200 C ABI -- the drop-in boundary of the B200-native walk-jump hot path.
 *
 * The reference (prescient-design/jamun) is pure Python and has no native boundary of its own;
 * every entry point below replaces one third-party kernel family the reference reaches through
 * Python (file:line under /root/reference/src/jamun are cited per function).  A maintainer binds
 * these with ctypes; jamun_b200/_lib.py is that binding.
 *
 * Conventions: the caller owns every buffer (device pointers unless noted); no entry point
 * allocates or synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 * the return value is 0 or a negative JAMUN_E* code (jamun_last_error() gives the text);
 * nothing throws.  Thread-safe for distinct streams.  All floating point is fp32, indices int32.
 *
 * Internal node-feature layout ("SoA irreps"): for irreps `S x0e + V x1e` a row is
 * [S scalars | V x-components | V y-components | V z-components]; jamun_layout_* convert from/to
 * the e3nn layout [S scalars | V x (x,y,z)] used at module boundaries of the reference.
 */
#ifndef JAMUN_B200_H
#define JAMUN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* jamun_stream_t;

#define JAMUN_OK 0
#define JAMUN_EINVAL (-1)   /* bad argument / unsupported irreps */
#define JAMUN_ECUDA (-2)    /* CUDA launch error */

/* hidden irreps the kernels are specialised for (e3conv.yaml:4): 120x0e + 32x1e */
#define JAMUN_S 120
#define JAMUN_V 32
#define JAMUN_HID (JAMUN_S + 3 * JAMUN_V)          /* 216 */
#define JAMUN_GATE_IN (JAMUN_S + JAMUN_V + 3 * JAMUN_V) /* 248 = 152x0e + 32x1e */
#define JAMUN_S0 56                                  /* initial embedding scalars */
#define JAMUN_EDGE_HID 64                            /* radial MLP hidden width (edge_attr_dim) */
#define JAMUN_NBASIS 32                              /* radial Gaussians */

int jamun_abi_version(void);
const char* jamun_last_error(void);

/* NoiseConditionalScaling.scale_predictor: out = W2 . SELU(W1*c + b1) + b2, optional sigmoid
 * (model/noise_conditioning.py:33-38,50-54,69-73).  w1,b1:[n]  w2:[n,n] row-major (out,in)  b2:[n]. */
int jamun_noise_mlp(const float* w1, const float* b1, const float* w2, const float* b2, float c_noise,
                    int n, int apply_sigmoid, float* out, jamun_stream_t stream);

/* AtomEmbeddingWithResidueInformation.forward fused with initial_noise_scaling
 * (model/atom_embedding.py:58-76, arch/e3conv.py:129-130).  idx*: [N] int32, tab*: [rows, dim*],
 * scale: [sum dims] or NULL, out: [N, sum dims]. */
int jamun_atom_embed(const int* idx0, const int* idx1, const int* idx2, const int* idx3,
                     const float* tab0, const float* tab1, const float* tab2, const float* tab3,
                     int dim0, int dim1, int dim2, int dim3, const float* scale, int N, float* out,
                     jamun_stream_t stream);

/* mean_center + input scaling (utils/mean_center.py:7-12, model/denoiser.py:205-207,191-192):
 * ybar = y - centroid_chain(y) (ybar = y when center == 0); p = c_in * ybar.  chain_ptr: [G+1].
 * ybar/p may be NULL. */
int jamun_center_scale(const float* y, const int* chain_ptr, int G, int center, float c_in, float* ybar, float* p,
                       jamun_stream_t stream);

/* radius_graph + bonded-edge concatenation as a receiver-sorted CSR
 * (model/denoiser.py:138-166 -> torch_geometric.nn.radius_graph -> torch_cluster radius kernel).
 * Pair (j->i) kept iff same chain, j != i, dx*dx+dy*dy+dz*dz < r2 (strict, unfused fp32), and j is
 * among the first max_num_neighbors+1 hits (self included) scanning the chain in ascending index;
 * max_num_neighbors < 0 disables the cap.  Row i = radial sources ascending, then the bonded
 * in-edges of i in their original order (bond_rowptr:[N+1], bond_src: CSR of the template's bonded
 * edge_index by receiver).  Outputs: rowptr:[N+1] (rowptr[N] = E on device), col/edst:[cap] source
 * and receiver per edge, ebond:[cap] 0 radial / 1 bonded.  cap >= (max_num_neighbors+1)*N + E_bonded.
 * scratch: [N+1] int32. */
int jamun_radius_csr(const float* pos, const int* chain_of, const int* chain_ptr, int N, float r2,
                     int max_num_neighbors, const int* bond_rowptr, const int* bond_src, int* scratch,
                     int* rowptr, int* col, int* edst, unsigned char* ebond, jamun_stream_t stream);

/* The same CSR from a cell-list search (north_star subsystem 1), for long chains: one CTA per chain bins the chain's atoms
 * into cells of edge >= r_cut in shared memory and tests the 27 surrounding cells with the same fp32 predicate; the cap rule
 * is applied as "the max_num_neighbors+1 smallest hit indices", which is what the ascending scan of torch_cluster's kernel
 * keeps.  nbr: [N, max_num_neighbors+1] int scratch; max_chain <= 8192 atoms; 0 <= max_num_neighbors < 64. */
int jamun_radius_csr_cells(const float* pos, const int* chain_ptr, int G, int N, int max_chain, float r2, float r_cut,
                           int max_num_neighbors, const int* bond_rowptr, const int* bond_src, int* scratch, int* nbr,
                           int* rowptr, int* col, int* edst, unsigned char* ebond, jamun_stream_t stream);

/* Edge featurisation (arch/e3conv.py:110-127): for each CSR edge, rhat = (p[src]-p[dst])/|.| (so
 * sh = [1, sqrt3*rhat]) and the 32 Gaussian radial bases exp(-((d-mu_k)/step)^2)/1.12.
 * rhat: [cap,4] (x,y,z,d)  rb: [cap, JAMUN_NBASIS]  mu: [JAMUN_NBASIS]. */
/* Out-edge lists: src_eid[src_rowptr[j] .. src_rowptr[j+1]) = positions e of the receiver-major edge list with
 * col[e] == j (order within a source unspecified).  scratch: [N+1] ints. */
int jamun_csr_by_source(const int* rowptr, const int* col, int N, int cap, int* scratch, int* src_rowptr, int* src_eid,
                        jamun_stream_t stream);

int jamun_edge_geom(const float* p, const int* rowptr, const int* col, const int* edst, int N, int cap,
                    const float* mu, float step, float* rhat, float* rb, jamun_stream_t stream);

/* Radial MLP hidden layer (e3tools/nn/_conv.py:84-94, _mlp.py:21-34), bondedness embedding folded:
 * h = SiLU(rb . w0r + b0eff[ebond]).  w0r: [32, 64] (the radial half of radial_nn.0.weight, transposed: basis-major),
 * b0eff: [2, 64] = bias + W0[:, :32] . embed_bondedness[flag].  h: [cap, 64]. */
int jamun_edge_radial_hidden(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap,
                             const float* w0r, const float* b0eff, float* h, jamun_stream_t stream);
/* All layers in one pass over the edges (the radial basis is read once): w0r_all [layers, 32, 64], b0eff_all
 * [layers, 2, 64], h_all [layers, cap, 64]. */
int jamun_edge_radial_hidden_all(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap,
                                 const float* w0r_all, const float* b0eff_all, int layers, float* h_all,
                                 jamun_stream_t stream);
/* The same on the warp-level tensor cores (mma.sync tf32, three-product split = fp32 accuracy).  jamun_radial_pack_frag turns
 * w0r_all [layers, 32, 64] into pre-split, fragment-ordered weight images img [layers * 4096] floats (once per plan);
 * jamun_edge_radial_hidden_mma evaluates all layers from them. */
int jamun_radial_pack_frag(const float* w0r_all, int layers, float* img, jamun_stream_t stream);
int jamun_edge_radial_hidden_mma(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap,
                                 const float* img, const float* b0eff_all, int layers, float* h_all, jamun_stream_t stream);

/* Conv.forward (e3tools/nn/_conv.py:96-119): gather, per-edge-weighted FullyConnectedTensorProduct,
 * scatter-mean -- evaluated in the aggregate-then-transform form: per receiver
 * A = sum_e [h_e,1] (x) f_e, out = alpha * A . M / max(1,deg).  x: [N, s_in + 3 v_in] SoA;
 * (s_in,v_in) in {(120,32),(56,0)}.  m0: [65, s_in+v_in, 152]  m1: [65, s_in+2 v_in, 32] packed from
 * radial_nn.3.{weight,bias}.  out: [N, 248] SoA (152 scalars | 32 x3). */
int jamun_conv_fwd(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                   const float* rhat, const float* m0, const float* m1, float alpha0, float alpha1, int N,
                   float* out, jamun_stream_t stream);

/* Tensor-core evaluation of Conv.forward, split in two launches (DESIGN.md "conv on tcgen05"):
 * (1) jamun_conv_build_a: per receiver, A[k',u'] = sum_e h'_e[k'] f_e[u'] written as the fp32 A operand of the GEMM in
 *     stage-major layout ([stage][rows_pad][32]; stage = k'*nslots + slot; slots per the table in conv_build.cu) for
 *     rows [row0, row0+nrows); a0 holds the 0e operand, a1 + c*a1_comp_stride the three 1e operands; inv_deg[i] = 1/max(1,deg).
 *     The path 0e(x)1e->1e is gathered from y[N, 2176] = x_s . W (pre-transformed source rows, 65*32 used columns) into p2[N, p2_ld]:
 *     p2[i, c*32+w] = sum_e rhat_e[c] sum_k' h'_e[k'] y[src_e, k'*32+w], multiplied by p2_scale/deg when p2_scale != 0.
 *     max_degree: an upper bound of the in-degree of every node (<= 64 selects the shared-memory-cached fast kernel).
 *     This is the FP32-pipe builder (exact fp32 aggregate); jamun_conv_build_tc + jamun_conv_p2 below is the default.
 * (2) jamun_gemm_tf32x3: out[r, out_col[s] + n] = row_scale[r] * alpha[s] * sum_K A_s[r,K] B_s[K,n] for up to 4 segments,
 *     tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-level accuracy), accumulators in TMEM.  b[s] is the weight operand
 *     pre-packed per stage as (hi | lo) images in the UMMA K-major SWIZZLE_128B shared-memory layout
 *     (jamun_b200/packing.py::pack_b_images).  rows_pad % 128 == 0; sum n_pad <= 256; n_pad % 16 == 0, <= 160.
 *     addend[s] (optional, [rows, addend_ld[s]]) is added to the accumulator before scaling.  col_blocks > 1 launches
 *     one CTA column per block of output columns: block y uses b[s] + y*b_block_floats and writes at out_col[s] + y*n_valid[s]
 *     (same A) -- used for the wide per-node transform Y = x_s . W (N = 65*32 = 2080, run as 17 blocks of 128; y row stride 2176).
 * (3) jamun_pack_rows: copies columns [col0, col0+ncols) of a row-major matrix into the GEMM's stage-major, chunk-swizzled
 *     A layout ([ceil(ncols/32)][rows_pad][32], zero padded). */
int jamun_conv_build_a(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                       const float* rhat, const float* y, int max_degree, int row0, int nrows, int rows_pad, float* a0, float* a1, long long a1_comp_stride,
                       float* p2, int p2_ld, float p2_scale, float* inv_deg, jamun_stream_t stream);
/* Tensor-core form of jamun_conv_build_a's aggregate (same a0 / a1 operands, 3xTF32 products accumulated in
 * tensor memory: elements agree with the FP32 builder to ~1e-6 relative).  Per receiver node the aggregate is the
 * small GEMM F^T.H over its in-edges (K <= 33), run as tcgen05.mma by a persistent warp-specialised kernel.
 * Writes inv_deg[i] = 1/max(1,deg) when inv_deg != NULL.  Does not compute the 0e(x)1e->1e gather: pair it with
 * jamun_conv_p2 (which may run concurrently on another stream). */
int jamun_conv_build_tc(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                        const float* rhat, int row0, int nrows, int rows_pad, float* a0, float* a1,
                        long long a1_comp_stride, float* inv_deg, jamun_stream_t stream);

/* Path 0e(x)1e->1e of the convolution, transform-then-aggregate:
 * p2[i, c*32+w] = sc * sum_{e->i} rhat_e[c] * T_e[w],  T_e[w] = sum_k' h'_e[k'] * y[col[e], k'*32+w],
 * sc = p2_scale/max(1,deg) (p2_scale != 0) or 1 (raw sums).  y: [N, 2176] rows from the per-node transform GEMM.
 * T is evaluated source-major (src_rowptr / src_eid from jamun_csr_by_source) into t_edge: [cap, 32] scratch.
 * Also writes inv_deg[i] = 1/max(1,deg) when inv_deg != NULL.  p2 == NULL: only t_edge is computed (the receiver-side sum
 * is then taken by jamun_tail_pack). */
int jamun_conv_p2(const int* rowptr, const int* src_rowptr, const int* src_eid, const float* h, const float* rhat,
                  const float* y, int N, float* t_edge, float* p2, int p2_ld, float p2_scale, float* inv_deg,
                  jamun_stream_t stream);

/* Weight re-layout on the device (replaces the host-side packing of jamun_b200/packing.py; same bytes): row-major fp32
 * weights -> the (hi | lo) stage images jamun_gemm_tf32x3 streams as its B operand, out[col_blocks][n_stages][2][n_pad*32]
 * (UMMA K-major SWIZZLE_128B; hi = w & 0xFFFFE000, lo = w - hi).  Element (k, col) comes from
 * src[(row + (col / n_inner) * outer_rows) * ld + col % n_inner] with row = row_map ? row_map[k] : k; rows outside
 * [0, K_src) and columns >= N_valid are zero.  transpose != 0 packs the transposed matrix instead (element (k, col) =
 * src[row_map ? row_map[col] : col][k], k < N_valid, col < K_src) -- the weight operand of the backward GEMMs.  The operands replaced are the re-laid-out radial_nn.3 weights of
 * e3tools/nn/_conv.py:84-94 and the o3.Linear weights of _interaction.py:23-24. */
int jamun_pack_b(const float* src, int ld, const int* row_map, int K_src, int n_stages, int N_valid, int n_inner,
                 int outer_rows, int n_pad, int col_blocks, int transpose, float* out, jamun_stream_t stream);

int jamun_pack_rows(const float* x, int ld, int col0, int ncols, int rows, int rows_pad, float* a, jamun_stream_t stream);
int jamun_gemm_tf32x3(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                      const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                      const int* addend_ld, int col_blocks, long long b_block_floats, int rows, int rows_pad,
                      const float* row_scale, float* out, int out_ld, jamun_stream_t stream);
/* Split-K form of jamun_gemm_tf32x3 for small row counts (few 128-row tiles would leave most SMs idle): k_splits CTAs per tile
 * each accumulate a contiguous share of every segment's stages and write a scaled partial result to
 * partial[k_splits, rows, out_ld] (scratch); a second kernel sums the partials in ascending order (deterministic).
 * No addend, no column blocks. */
int jamun_gemm_tf32x3_splitk(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                             const int* n_valid, const int* out_col, const float* alpha, int rows, int rows_pad,
                             const float* row_scale, float* out, int out_ld, int k_splits, float* partial,
                             jamun_stream_t stream);

/* fp16-split form of the same GEMM (the default of the sampling path; csrc/gemm_tf32x3.cu): the operands are split as
 * hi = rn16(v), lo = rn16(v - hi) -- fp16 and tf32 both carry an 11-bit significand, so a_hi.b_hi + a_hi.b_lo + a_lo.b_hi has
 * the accuracy of the 3xTF32 form -- and contracted with tcgen05.mma kind::f16 (K = 16 per instruction, twice the tf32 rate).
 * a: the same fp32 stage-major operand; b: images from jamun_pack_b_f16 (n_pad rows x 128 bytes [hi k0..31 | lo k0..31] per
 * stage, K-major SWIZZLE_128B; b_block_floats counts 4-byte words).  The weights are pre-scaled by a power of two when packed
 * (fp16 range), the caller folds 1/scale into alpha and passes the scale as addend_scale[s] (the addend joins the scaled
 * accumulator: out = (acc + addend_scale * addend) * alpha * row_scale; NULL: 1).  status (or NULL): bit 0 is OR-ed in when an element of a exceeds the fp16
 * range (the result is then invalid and the caller must use jamun_gemm_tf32x3).  k_splits > 1: split-K as
 * jamun_gemm_tf32x3_splitk (partial scratch, no addend, no column blocks).  a_tile_major: the operand layout of
 * jamun_conv_build_tc_tiled. */
int jamun_pack_b_f16(const float* src, int ld, const int* row_map, int K_src, int n_stages, int N_valid, int n_inner,
                     int outer_rows, int n_pad, int col_blocks, int transpose, float scale, float* out, jamun_stream_t stream);
int jamun_gemm_f16x3(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                     const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                     const int* addend_ld, const float* addend_scale, int col_blocks, long long b_block_floats, int rows,
                     int rows_pad, const float* row_scale, float* out, int out_ld, int k_splits, float* partial, int* status,
                     int a_tile_major, jamun_stream_t stream);
/* Fused ConvBlock epilogues of jamun_gemm_f16x3 (the accumulators never leave the SM un-activated):
 * mode 1 (contraction, segments [160 | 32 | 32 | 32]): Gate (e3tools/nn/_gate.py:63) -> the block-tail operands
 *        op_s stages 0-3 (activated scalars, zero padded to 128) and op_v[c] stage 0 (gated vectors); addend[1..3] carries the
 *        receiver-side sums of the 0e(x)1e->1e path (jamun_conv_p2 with p2_scale = 0).  Replaces jamun_tail_pack.
 * mode 2 (block tail, segments [128 | 32 | 32 | 32], out_col [0,120,152,184]): x_new = skip_w ? x_res*w + y*(1-w) : y and
 *        x_scaled = x_new * s_next (model/noise_conditioning.py:50-73, arch/e3conv.py:131-133; weights per irrep [120 | 32]),
 *        both row-major [rows, 216], plus x_scaled packed as the x_in halves of the next block's operands: op_s stages 4-7
 *        (also the per-node transform's operand) and op_v[c] stage 1.  x_scaled / op_s may be NULL (last block).  Replaces
 *        jamun_tail_mix.
 * op_s: [8][op_rows_pad][32], op_v: 3 x [2][op_rows_pad][32] (component stride op_v_comp_stride floats), stage-major, 16-byte
 * chunks XOR-ed with row & 7 -- the A layout of this GEMM. */
typedef struct {
    int mode;
    float* op_s;
    float* op_v;
    long long op_v_comp_stride;
    int op_rows_pad;
    float c_act, c_gate;   /* mode 1: normalize2mom constants of the Gate */
    const float* x_res;    /* mode 2 */
    const float* skip_w;
    const float* s_next;
    float* x_new;
    float* x_scaled;
} jamun_gemm_epilogue;
int jamun_gemm_f16x3_fused(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                           const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                           const int* addend_ld, const float* addend_scale, int rows, int rows_pad, const float* row_scale,
                           int* status, int a_tile_major, const jamun_gemm_epilogue* epi, jamun_stream_t stream);
/* jamun_conv_build_tc writing the tile-major operand layout [row / 128][stage][128][32] (all stages of a 128-row tile
 * contiguous; rows_pad % 128 == 0), consumed by jamun_gemm_f16x3 with a_tile_major = 1: the 715 lines a node contributes then
 * fall into one ~12 MB region instead of being rows_pad * 128 bytes apart, and a GEMM CTA streams its tile sequentially. */
int jamun_conv_build_tc_tiled(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                              const float* rhat, int row0, int nrows, int rows_pad, float* a0, float* a1,
                              long long a1_comp_stride, float* inv_deg, jamun_stream_t stream);

/* Gate + self-interaction + skip Linear + noise-conditional skip/scale
 * (e3tools/nn/_gate.py:63-64, _interaction.py:26-30, model/noise_conditioning.py:50-73,
 * arch/e3conv.py:131-133).  y = Lin_self(Gate(conv)) + Lin_skip(x_in);
 * x_new = mix ? x_res*w + y*(1-w) : y;  x_scaled = x_new * s_next (if s_next).
 * Weights pre-scaled by 1/sqrt(fan_in): wself_s:[120,120] wself_v:[32,32] wskip_s:[s_in,120]
 * wskip_v:[32,32] or NULL.  skip_w, s_next: [152] per-irrep or NULL.
 * vadd: [N, 96] or NULL -- added to the 1e part of conv before the gate (the 0e(x)1e->1e path when it was
 * gathered by jamun_conv_p2 on another stream instead of through the GEMM epilogue). */
int jamun_block_tail(const float* conv, const float* vadd, const float* x_in, int s_in, int v_in, const float* x_res,
                     const float* wself_s, const float* wself_v, const float* wskip_s, const float* wskip_v,
                     const float* skip_w, const float* s_next, float c_act, float c_gate, int N,
                     float* x_new, float* x_scaled, jamun_stream_t stream);

/* The same block tail as three launches with the two Linears on the tensor cores:
 * jamun_tail_pack  Gate(conv (+vadd)) and x_in -> stage-major operands for jamun_gemm_tf32x3:
 *                  a_s: [4 + ceil(s_in/32)][rows_pad][32] (activated scalars | input scalars, zero padded per 32),
 *                  a_v: 3 components x [1 + (v_in>0)][rows_pad][32] (gated vectors | input vectors), stride a_v_comp_stride;
 *                  t_edge != NULL: adds p2_scale/deg * sum_{e->i} rhat_e[c] t_edge[e, w] to the 1e part (jamun_conv_p2 with
 *                  p2 == NULL); conv_has_v == 0: the 1e columns of conv are not read (initial block);
 * jamun_gemm_tf32x3 with B = [W_self ; W_skip] images (K zero-padded per 32) -> y [N, 216];
 * jamun_tail_mix   x_new = skip_w ? x_res*w + y*(1-w) : y;  x_scaled = x_new * s_next;  xs_op (or NULL): the 120 scalars of
 *                  x_scaled as the [4][rows_pad][32] operand of the next block's per-node transform (= jamun_pack_rows;
 *                  positions 120..127 of the last stage are not written: keep them zero). */
int jamun_tail_pack(const float* conv, const float* vadd, const float* x_in, int s_in, int v_in, float c_act,
                    float c_gate, int N, int rows_pad, float* a_s, float* a_v, long long a_v_comp_stride,
                    const int* rowptr, const float* rhat, const float* t_edge, float p2_scale, int conv_has_v,
                    jamun_stream_t stream);
int jamun_tail_mix(const float* y, const float* x_res, const float* skip_w, const float* s_next, int N, float* x_new,
                   float* x_scaled, float* xs_op, int rows_pad, jamun_stream_t stream);

/* Output head (e3tools/nn/_mlp.py:37-114, arch/e3conv.py:134-135): Linear -> Gate -> Linear(1x1e) * gain.
 * w1_s: [120,152] w1_v:[32,32] pre-scaled; w2: [32] pre-scaled by gain/sqrt(32).  g: [N,3]. */
int jamun_head(const float* x, const float* w1_s, const float* w1_v, const float* w2, float c_gate, int N,
               float* g, jamun_stream_t stream);

/* One fused walk-jump step (sampling/mcmc/functional/_splitting.py:26-41,136-178, model/denoiser.py:111-114,
 * 200,213-215, sampling/walkjump/_single_measurement.py:57): given the network output g at the current y,
 *   xhat = center(c_skip*ybar + c_out*g); score = (xhat - y)/sigma^2; psi = beta*clip(score)
 *   [save y/xhat/score]; v = first ? v : v + (delta/2) psi            (closing B of the previous step)
 *   if !last: v += u (delta/2) psi; y += (delta/2) v; v = a v + z sqrt(u) R; y += (delta/2) v   (B A O A)
 *             ybar = center(y); p = c_in*ybar                          (prologue of the next evaluation)
 * R = noise[N,3] if non-NULL else Philox4x32-10(seed, step) Box-Muller.  clip<=0 disables clipping.
 * traj_* may be NULL.  score_in non-NULL: use that score instead of deriving xhat/score from g (generic
 * score_fn protocol of mcmc/_splitting.py:56-58; g, xhat outputs are then ignored and may be NULL).
 * dev_state non-NULL (CUDA-graph replay): the Philox step is dev_state[0] and the trajectory frame written is
 * traj_* + dev_state[1] * 3 * n_atoms; jamun_walk_advance(dev_state, slot_inc) bumps both between replays. */
typedef struct {
    float c_in, c_skip, c_out, sigma2;
    float delta, u, a, z_sqrt_u, beta, clip;
    int first, last, center;   /* center: 0 disables both mean-centrings (Denoiser(mean_center=False)) */
    unsigned long long seed, step;
} jamun_walk_params;

int jamun_walk_step(float* y, float* v, float* ybar, float* p, const float* g, const float* score_in,
                    const int* chain_ptr, int G, const jamun_walk_params* prm, const float* noise, float* xhat, float* score,
                    float* traj_y, float* traj_xhat, float* traj_score, const unsigned long long* dev_state,
                    jamun_stream_t stream);
int jamun_walk_advance(unsigned long long* dev_state, int slot_inc, jamun_stream_t stream);

/* out = a*x + b*noise with noise = given or Philox (initial y = x + sigma*eps, v0 = sqrt(u)*eps;
 * utils/sampling_wrapper.py:21-24, functional/_splitting.py:11-23). */
/* ABOBA pieces (functional/_splitting.py:44-109): drift y += (delta/2) v; and, given the score at the half
 * step, psi = beta*clip(score); v += u (delta/2) psi; v = a v + z sqrt(u) R; v += (delta/2) psi; y += (delta/2) v. */
int jamun_aboba_drift(float* y, const float* v, float half_delta, int n_atoms, jamun_stream_t stream);
int jamun_aboba_kick(float* y, float* v, const float* score, const jamun_walk_params* prm, const float* noise,
                     int n_atoms, jamun_stream_t stream);

int jamun_gaussian_axpy(const float* x, float a, float b, const float* noise, unsigned long long seed,
                        unsigned long long step, int n_atoms, float* out, jamun_stream_t stream);

/* Dense layer used by the module-level compatibility forwards (ScalarMLP of e3tools/nn/_mlp.py:10-34 on arbitrary
 * edge_attr): out[r, o] = act(sum_k w[o, k] in[r, k] + b[o]); act 0 = identity, 1 = SiLU. */
int jamun_linear_act(const float* in, const float* w, const float* b, int rows, int K, int O, int act, float* out,
                     jamun_stream_t stream);

/* e3nn.o3.FullyConnectedTensorProduct(in1, in2, out, shared_weights=False, internal_weights=False) for l <= 1 with per-row
 * weights: the reference's plug-in seam `tp(x_src, sh, weight)` (e3tools/nn/_conv.py:76-94).  x1:[Z,d1] x2:[Z,d2] w:[Z,w_ld]
 * in e3nn layouts; instr: [n_instr][12] int32 records (off1, mul1, l1, off2, mul2, l2, off_out, mul_out, l_out, weight offset,
 * path coefficient sqrt(dim_out/fan_in) as float bits, 0) in e3nn's instruction order; out:[Z,d_out].  Not on the sampling
 * path (which never materialises per-edge weights). */
int jamun_tensor_product(const float* x1, int d1, const float* x2, int d2, const float* w, long long w_ld, const int* instr,
                         int n_instr, int d_out, int Z, float* out, jamun_stream_t stream);

/* e3nn layout <-> SoA layout for `s x0e + v x1e` rows. */
int jamun_layout_to_soa(const float* in, int s, int v, int N, float* out, jamun_stream_t stream);
int jamun_layout_from_soa(const float* in, int s, int v, int N, float* out, jamun_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training path: backward of the network, loss and alignment.  Reference: model/denoiser.py:87-109,
 * 219-319 (noise, align, loss, training_step), utils/align.py:9-56 (Kabsch); the backward of e3nn / torch_scatter ops the
 * reference obtains from autograd is written out here per forward kernel.  All reductions have a fixed order (no atomics).
 * ------------------------------------------------------------------------------------------------------------------ */

/* Y[n, :b] (+)= X[n, :a] . W, W row-major [a, b] (transW: [b, a], i.e. Y = X . W^T).  b <= 160.  rows_dev (optional): device
 * row count (clamped to rows).  Forward / recompute of the o3.Linear stand-ins (_interaction.py:23-24) and their dX. */
int jamun_rowmat_mul(const float* X, int ldx, const float* W, int ldw, int transW, float* Y, int ldy, int rows,
                     const int* rows_dev, int a, int b, int accumulate, jamun_stream_t stream);
/* dW (+)= X^T . dY (a reduction over nodes or edges in row chunks, partial sums added in ascending order).
 * scratch: jamun_rowmat_dw_scratch(rows, a, b) floats. */
long long jamun_rowmat_dw_scratch(int rows, int a, int b);
int jamun_rowmat_dw(const float* X, int ldx, const float* dY, int ldy, float* dW, int lddw, int transW, int rows,
                    const int* rows_dev, int a, int b, int accumulate, float* scratch, jamun_stream_t stream);
/* out[c] (+)= sum_n M[n, c] (rows with flag[n] == flag_value when flag != NULL); fold_v > 0 folds SoA columns
 * [fold_s | fold_v x3] to per-irrep [fold_s + fold_v].  scratch: jamun_colsum_scratch(rows, cols) floats. */
long long jamun_colsum_scratch(int rows, int cols);
int jamun_colsum(const float* M, int ld, int rows, const int* rows_dev, int cols, const unsigned char* flag, int flag_value,
                 int fold_s, int fold_v, float* out, int accumulate, float* scratch, jamun_stream_t stream);

/* Conv backward (e3tools/nn/_conv.py:93-119): see jamun_b200/csrc/train_conv_bwd.cu for the scheme.
 * jamun_conv_bwd_scale   g = dOut * inv_deg * (alpha0 | alpha1)                                   [N, 248]
 * jamun_stage_atb        acc[u, w] = sum_c sum_r A_c[stage][r][u] * B[r][b_col0 + c*b_comp_stride + w] for every stage of a
 *                        stage-major GEMM operand (dM = A^T . G, dM2 = dY^T . x_s, tail/weight gradients); mode 0 scatters
 *                        operand slots to output rows (slot_row0/slot_rows: HOST arrays of nslots entries), mode 1 writes
 *                        the transpose out[(k*out_rows + w)*32 + u].
 * jamun_conv_bwd_edge    per receiver: dh[e, :64] and dxe[e, :D_in] from dA0 [N, ld0] / dA1 [3][N, ld1] (rows of G . M^T).
 * jamun_conv_bwd_p2      path 0e(x)1e->1e, source-major: dh[e] += Y_j . dT_e, dY operand [65][rows_pad][32].
 * jamun_conv_bwd_gather  dx[j] = sum_{e in out(j)} dxe[e] (+ extra[j, :n_extra]), ascending edge id. */
int jamun_conv_bwd_scale(const float* dout, const float* inv_deg, float alpha0, float alpha1, int N, float* g,
                         jamun_stream_t stream);
int jamun_stage_atb(const float* a, long long a_comp_stride, int ncomp, int n_stages, int nslots, int rows, int rows_pad,
                    const float* b, int ldb, int b_col0, int b_comp_stride, int W, float* out, int mode, int out_rows,
                    const int* slot_row0, const int* slot_rows, jamun_stream_t stream);
/* Tensor-core form of jamun_stage_atb (tcgen05 kind::tf32, 3xTF32, TS mode; both operands MN-major, see csrc/gemm_atb.cu):
 * jamun_pack_rows_split writes the (hi | lo) images [2][nslots][rows_pad][32] of x[:, col0:col0+ncols] (one call per component,
 * images consecutive); jamun_stage_atb_tc consumes them.  The node range is split over k_splits CTAs per 128-row tile and the
 * partial tiles are summed in ascending order.  partial: jamun_stage_atb_tc_scratch(n_stages, W, k_splits) floats. */
int jamun_pack_rows_split(const float* x, int ld, int col0, int ncols, int rows, int rows_pad, int nslots, float* out,
                          jamun_stream_t stream);
long long jamun_stage_atb_tc_scratch(int n_stages, int W, int k_splits);
int jamun_stage_atb_tc(const float* a, long long a_comp_stride, int ncomp, int n_stages, int nslots, int rows, int rows_pad,
                       const float* bsplit, int W, float* out, int mode, int out_rows, const int* slot_row0,
                       const int* slot_rows, int k_splits, float* partial, jamun_stream_t stream);
int jamun_conv_bwd_edge(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                        const float* rhat, const float* dA0, int ld0, const float* dA1, int ld1, long long dA1_comp_stride,
                        int N, float* dh, float* dxe, jamun_stream_t stream);
int jamun_conv_bwd_p2(const int* src_rowptr, const int* src_eid, const int* edst, const float* h, const float* rhat,
                      const float* y, int y_ld, const float* g, int N, int rows_pad, float* dh, float* dy_op,
                      jamun_stream_t stream);
int jamun_conv_bwd_gather(const int* src_rowptr, const int* src_eid, const float* dxe, int D, const float* extra,
                          int extra_ld, int n_extra, int N, float* dx, jamun_stream_t stream);

/* Gate (e3tools/nn/_gate.py:63-64), SoA layout: gated [N, 216] from conv [N, 248] and its backward. */
int jamun_gate_fwd(const float* conv, float c_act, float c_gate, int N, float* gated, jamun_stream_t stream);
int jamun_gate_bwd(const float* conv, const float* dgated, float c_act, float c_gate, int N, float* dconv,
                   jamun_stream_t stream);
/* Backward of x_new = skip_w ? x_res*w + y*(1-w) : y; x_scaled = x_new*s (arch/e3conv.py:132-133):
 * dy, dx_res and the per-element products whose column sums are ds (prod_s) and dw (prod_w). */
int jamun_mix_bwd(const float* dx_new, const float* dx_scaled, const float* y, const float* x_res, const float* skip_w,
                  const float* s_next, int N, float* dy, float* dx_res, float* prod_s, float* prod_w, jamun_stream_t stream);
/* Output head backward, elementwise part (e3tools/nn/_mlp.py:37-114): pre [N,32] gate pre-activations, hv [N,3,32]. */
int jamun_head_bwd(const float* pre, const float* hv, const float* w2, const float* dg, float c_gate, int N, float* dpre,
                   float* dhv, float* prod_w2, jamun_stream_t stream);
/* Radial MLP hidden layer backward: dz = dh * SiLU'(rb . w0r + b0eff[ebond])  [cap, 64] (rows < rowptr[N]). */
int jamun_radial_bwd(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap, const float* w0r,
                     const float* b0eff, const float* dh, float* dz, jamun_stream_t stream);
/* Atom embedding backward (model/atom_embedding.py:58-76): dtab_k[r] = scale * sum_{i: idx_k[i] = r} dx0[i, cols_k];
 * prod [N, D] = dx0 * embedding (column sums = dscale). */
int jamun_embed_bwd(const int* idx0, const int* idx1, const int* idx2, const int* idx3, const float* tab0, const float* tab1,
                    const float* tab2, const float* tab3, int dim0, int dim1, int dim2, int dim3, int rows0, int rows1,
                    int rows2, int rows3, const float* scale, const float* dx0, int N, float* dtab0, float* dtab1,
                    float* dtab2, float* dtab3, float* prod, jamun_stream_t stream);
/* NoiseConditionalScaling MLP backward (model/noise_conditioning.py:33-38). */
int jamun_noise_mlp_bwd(const float* w1, const float* b1, const float* w2, const float* b2, float c_noise, int n,
                        int apply_sigmoid, const float* dout, float* dw1, float* db1, float* dw2, float* db2,
                        jamun_stream_t stream);
/* xhat = center_chain(c_skip*ybar + c_mix*g) (model/denoiser.py:200,213-215; ybar may be NULL).  Its backward with respect
 * to g is the same map applied to dxhat with ybar = NULL. */
int jamun_combine_xhat(const float* g, const float* ybar, const int* chain_ptr, int G, float c_skip, float c_mix, int center,
                       float* xhat, jamun_stream_t stream);
/* Coordinate loss (model/denoiser.py:251-287) per chain: raw = mean |xhat - x|^2, loss = raw * loss_weight * scale
 * (scale = 1/c_out^2), rmsd = mean |xhat - x| / (sigma sqrt3); backward: dxhat = dloss * 2 (xhat - x) loss_weight scale / n. */
int jamun_loss_fwd(const float* xhat, const float* x, const int* chain_ptr, int G, float scale, float sigma,
                   const float* loss_weight, float* loss, float* raw, float* rmsd, jamun_stream_t stream);
int jamun_loss_bwd(const float* xhat, const float* x, const int* chain_of, const int* chain_ptr, int N, float scale,
                   const float* loss_weight, const float* dloss, float* dxhat, jamun_stream_t stream);
/* Batched Kabsch alignment (utils/align.py:9-56): out = R y + t per chain, R = V diag(1,1,det) U^T from the SVD of
 * H = sum y_c x_c^T (one-sided Jacobi in fp64).  rot (optional): [G, 9] rotation matrices. */
int jamun_kabsch_align(const float* y, const float* x, const int* chain_ptr, int G, float* out, float* rot,
                       jamun_stream_t stream);
/* Per chain: (sum of |x_i - x_j|^2 over pairs i > j with |x_i - x_j| < cutoff, number of such pairs) as two doubles
 * (utils/average_squared_distance.py:154-177; cutoff <= 0: all pairs).  sums: [G, 2] double. */
int jamun_avg_sq_dist(const float* pos, const int* chain_ptr, int G, float cutoff, double* sums, jamun_stream_t stream);
/* ema = decay*ema + (1-decay)*p over a flat buffer (callbacks/_ema.py, EMAOptimizer). */
int jamun_ema_update(float* ema, const float* p, float decay, long long n, jamun_stream_t stream);
/* out[:, col0 : col0+n] += add[:, :n]. */
int jamun_add_cols(float* out, int ld, int col0, const float* add, int add_ld, int n, int N, jamun_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
