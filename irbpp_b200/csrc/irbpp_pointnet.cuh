// irbpp_pointnet.cuh -- the next item's point-cloud feature on the device.
//
// Reference (model.py:328-335, :366-372): every forward pass gathers shapeArray[next_item_ID.cpu()] on the HOST
// ([B, 100000, 3] float32 rows), draws ONE index set  np.random.randint(P, size=samplePointsNum)  shared by the
// whole batch, copies [B, 1024, 3] to the device and runs  shapeEncoder = Linear(3,128) -> LeakyReLU ->
// Linear(128,128) -> LeakyReLU  (model.py:266-270) followed by a max over the points.
//
// Here shapeArray stays resident in HBM, the index set comes from a counter-based generator (same role: one
// uniform-with-replacement index set per forward pass), and -- because the index set is shared by the batch --
// the encoded feature depends only on the SHAPE, not on the bin: the encoder runs once per library shape
// (S x 1024 points) instead of once per bin (B x 1024 points), a 4096 / S-fold cut of the work at B = 4096, and
// the per-bin result is a gather of 128 floats.  Two entry points (irbpp.cu): the drop-in gather of the sampled
// clouds [B, n, 3] for callers that keep their own encoder, and the fused feature path.
// Float32 arithmetic with FMA contraction disabled (the library is built -fmad=false); the sums of the two
// layers are accumulated in a fixed order, which differs from cuBLAS's: tests state the tolerance.
#pragma once
#include <stdint.h>

namespace irbpp {

constexpr int PN_H = 128;          // width of both shapeEncoder layers (model.py:267,269)
constexpr int PN_TILE = 64;        // points per CTA
constexpr int PN_THREADS = 128;
constexpr int PN_W2_LD = PN_H + 4; // padded row of the transposed second-layer weight in shared memory

struct PointNetParams {
    const float* shape_array;      // [S][P][3] device
    int32_t S, P, n_points;
    uint64_t seed, counter;        // index j of this forward pass: mix(seed, counter, j) mod P
    const float* W1; const float* b1;    // [128][3], [128]   (nn.Linear layout: out x in)
    const float* W2; const float* b2;    // [128][128], [128]
    float slope;                   // LeakyReLU negative slope (0.01)
    int32_t* feat_keys;            // [S][128] running maxima as order-preserving integer keys
    // gather
    const float* obs; int64_t obs_stride; int32_t item_col;   // item id = (int) obs[b * obs_stride + item_col] ...
    const int32_t* ids;            // ... or ids[b] when not NULL
    int32_t B;
    float* out;                    // features [B][128] or clouds [B][n_points][3]
    int32_t* indices_out;          // [n_points] or NULL
};

__host__ __device__ __forceinline__ uint32_t pn_index(uint64_t seed, uint64_t counter, uint32_t j, uint32_t P) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((counter << 20) ^ (uint64_t)j) + 0xD1B54A32D192ED03ull;      // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)((z >> 32) % (uint64_t)P);
}

__device__ __forceinline__ int pn_key(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float pn_unkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
__device__ __forceinline__ float pn_leaky(float v, float slope) { return v > 0.0f ? v : v * slope; }

__global__ void irbpp_pn_init_kernel(int32_t* keys, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) keys[i] = (int32_t)0x80000000;
}

// One CTA: PN_TILE sampled points of one shape through both layers, column maxima into feat_keys[shape].
__global__ void __launch_bounds__(PN_THREADS) irbpp_shape_encode_kernel(const PointNetParams Q) {
    extern __shared__ __align__(16) float pn_smem[];
    float* W2t = pn_smem;                                  // [128 k][PN_W2_LD]: W2t[k][c] = W2[c][k]
    float* h1 = W2t + PN_H * PN_W2_LD;                     // [PN_TILE][128]
    float* pts = h1 + PN_TILE * PN_H;                      // [PN_TILE][4]
    float* red = pts + PN_TILE * 4;                        // [8][128] partial column maxima
    const int tid = threadIdx.x;
    const int tiles = (Q.n_points + PN_TILE - 1) / PN_TILE;
    const int shape = blockIdx.x / tiles, tile = blockIdx.x - shape * tiles;
    const int p0 = tile * PN_TILE;
    const int np = min(PN_TILE, Q.n_points - p0);
    for (int i = tid; i < PN_H * PN_H; i += PN_THREADS) {  // coalesced read of W2[c][k], transposed store
        const int c = i >> 7, k = i & 127;
        W2t[k * PN_W2_LD + c] = Q.W2[i];
    }
    if (tid < PN_TILE) {
        float x = 0.0f, y = 0.0f, z = 0.0f;
        if (tid < np) {
            const uint32_t idx = pn_index(Q.seed, Q.counter, (uint32_t)(p0 + tid), (uint32_t)Q.P);
            const float* src = Q.shape_array + ((int64_t)shape * Q.P + idx) * 3;
            x = src[0]; y = src[1]; z = src[2];
        }
        pts[tid * 4 + 0] = x; pts[tid * 4 + 1] = y; pts[tid * 4 + 2] = z;
    }
    __syncthreads();
    {   // layer 1: thread = output channel
        const float w0 = Q.W1[tid * 3 + 0], w1 = Q.W1[tid * 3 + 1], w2 = Q.W1[tid * 3 + 2], bb = Q.b1[tid];
        for (int p = 0; p < PN_TILE; ++p) {
            const float v = ((pts[p * 4] * w0 + pts[p * 4 + 1] * w1) + pts[p * 4 + 2] * w2) + bb;
            h1[p * PN_H + tid] = pn_leaky(v, Q.slope);
        }
    }
    __syncthreads();
    // layer 2: 8 x 8 register tile per thread (rows = points ty*8.., columns = channels tx*8..)
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    for (int k = 0; k < PN_H; ++k) {
        float a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = h1[(ty * 8 + i) * PN_H + k];
        const float4 b0 = *reinterpret_cast<const float4*>(W2t + k * PN_W2_LD + tx * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(W2t + k * PN_W2_LD + tx * 8 + 4);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);      // explicit FMA (the library is built -fmad=false)
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float bb = Q.b2[tx * 8 + j];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (ty * 8 + i < np) { const float v = pn_leaky(acc[i][j] + bb, Q.slope); m = v > m ? v : m; }
        red[ty * PN_H + tx * 8 + j] = m;
    }
    __syncthreads();
    {
        float m = red[tid];
#pragma unroll
        for (int g = 1; g < 8; ++g) { const float v = red[g * PN_H + tid]; m = v > m ? v : m; }
        if (np > 0) atomicMax(Q.feat_keys + (int64_t)shape * PN_H + tid, pn_key(m));
    }
}

__device__ __forceinline__ int pn_item_of(const PointNetParams& Q, int b) {
    int id = Q.ids ? Q.ids[b] : (int)Q.obs[(int64_t)b * Q.obs_stride + Q.item_col];
    return id < 0 ? 0 : (id >= Q.S ? Q.S - 1 : id);
}

// features of every bin: the row of its next item
__global__ void irbpp_feature_gather_kernel(const PointNetParams Q) {
    const int64_t n = (int64_t)Q.B * PN_H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i >> 7), c = (int)(i & 127);
        Q.out[i] = pn_unkey(Q.feat_keys[(int64_t)pn_item_of(Q, b) * PN_H + c]);
    }
}

// drop-in for model.py:330-332: nextShape[b, j, :] = shapeArray[item_b, indices[j], :]
__global__ void irbpp_cloud_gather_kernel(const PointNetParams Q) {
    const int64_t n = (int64_t)Q.B * Q.n_points;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / Q.n_points), j = (int)(i - (int64_t)b * Q.n_points);
        const uint32_t idx = pn_index(Q.seed, Q.counter, (uint32_t)j, (uint32_t)Q.P);
        if (b == 0 && Q.indices_out) Q.indices_out[j] = (int32_t)idx;
        const float* src = Q.shape_array + ((int64_t)pn_item_of(Q, b) * Q.P + idx) * 3;
        float* dst = Q.out + i * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

constexpr int PN_SMEM_BYTES = (PN_H * PN_W2_LD + PN_TILE * PN_H + PN_TILE * 4 + 8 * PN_H) * (int)sizeof(float);

}  // namespace irbpp
