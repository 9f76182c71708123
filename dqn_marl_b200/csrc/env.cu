// Batched Louvre_Evacuation environment — fused reset / step kernels (sm_100a).
//
// A GROUP of WPE warps steps one env instance: EvacuationEnv.step of the reference
// (Louvre_Evacuation/envs/evacuation_env.py:122-172) = move_robot (map.py:160-202) + People.run
// (people.py:196-253) + fire update (fire_model.py:63-67) + _calculate_reward (:174-288) + done (:155-157)
// + _get_state (:84-120) in ONE launch.  WPE = 1 (one warp per env, 8 envs per CTA, only __syncwarp between
// phases) for small envs such as the 150-people room; WPE = 8 (one CTA per env) for the 1000-people grids.
// The occupancy map of the env (People.rmap, 1 bit per cell) is staged in shared memory for the whole step;
// person state streams through registers, SoA across envs in HBM.  Everything that is fp64 in the reference
// stays fp64 and is evaluated in the reference's operation order (compiled with -fmad=false), so
// rewards / health / accumulators are bit-identical.
//
// Sequential semantics reproduced in parallel:
//   * proposals only read rmap as it was before phase 4 (people.py:211-230) -> embarrassingly parallel; movers
//     are compacted and the 8 directions of a mover are scored by 4 lanes (2 directions each), reduced with the
//     reference's "first strictly greater wins" rule;
//   * move_plan is a dict keyed by target cell in first-proposer order (people.py:228-230): the key of a
//     target is min(list index of its proposers) (shared-memory hash table + atomicMin);
//   * random.shuffle picks the mover (people.py:239): keyed priority, atomicMin on (prio<<32 | index);
//   * rmap is a FLAG map written in key order (people.py:301-302,312-314): the final bit of a cell is the
//     write with the largest key, i.e. SET (or CLEAR when the cell evacuates) iff key(cell) > max key of
//     the winners that left it; a left cell that is nobody's target is simply cleared.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "common.h"
#include "philox.cuh"

namespace mq {

// smallest group (warps per env) that sets its last warp aside for the sequential health sum
#ifndef MQ_CHAIN_MIN_WPE
#define MQ_CHAIN_MIN_WPE 8
#endif
constexpr int MAXR = MQ_MAX_ROBOTS;
constexpr uint32_t HEMPTY = 0xFFFFFFFFu;

struct DevLayout {
    int L, W, stride, G, wpr, rmap_words;
    int n_fire_steps;
    int ctr_box[4], int_box[4];
    int robot_range[2];
    int robot_start[MAXR][2];
    int reset_center[2];
    int obs_exit[2];
    const double* dp5;
    const uint8_t* cellinfo;
    const double* danger_ctr;
    const double* danger_int;
    // several layouts in one batch (mq_env_create_layouts): tables of layout k start at dp5 + k*G*8, cellinfo + k*G and at the
    // danger-table offsets of its LMETA row; env_layout[env] = k.  All layouts share L x W and the number of fire steps.
    const int* env_layout;
    const int* meta;
};

// per-layout ints (row of DevLayout::meta)
enum { LM_CTR_BOX = 0, LM_INT_BOX = 4, LM_ROBOT_RANGE = 8, LM_ROBOT_START = 10, LM_RESET_CENTER = 18, LM_OBS_EXIT = 20,
       LM_CTR_OFF = 22 /* long long, in doubles */, LM_INT_OFF = 24, LMETA = 32 };

// The kernels take the layout of a single-layout batch straight from the kernel parameters (constant bank).  MULTI variants
// overwrite their private copy with the tables of this env's layout.
template <bool MULTI>
__device__ __forceinline__ void select_layout(DevLayout& lay, int env) {
    if (!MULTI) return;
    const int li = __ldg(lay.env_layout + env);
    const int* m = lay.meta + (size_t)li * LMETA;
#pragma unroll
    for (int k = 0; k < 4; ++k) { lay.ctr_box[k] = __ldg(m + LM_CTR_BOX + k); lay.int_box[k] = __ldg(m + LM_INT_BOX + k); }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        lay.robot_range[k] = __ldg(m + LM_ROBOT_RANGE + k);
        lay.reset_center[k] = __ldg(m + LM_RESET_CENTER + k);
        lay.obs_exit[k] = __ldg(m + LM_OBS_EXIT + k);
    }
#pragma unroll
    for (int k = 0; k < MAXR * 2; ++k) lay.robot_start[k >> 1][k & 1] = __ldg(m + LM_ROBOT_START + k);
    lay.dp5 += (size_t)li * lay.G * 8;
    lay.cellinfo += (size_t)li * lay.G;
    lay.danger_ctr += *reinterpret_cast<const long long*>(m + LM_CTR_OFF);
    lay.danger_int += *reinterpret_cast<const long long*>(m + LM_INT_OFF);
}

struct DevCfg {
    int n_envs, N, n_pad, R;
    unsigned long long seed;
    PhiloxKeys pk;                     // round keys of `seed`
    int env_id_base, max_steps, reset_robots, reset_fire, auto_reset;
    int hash_cap, n_leaf_max;          // proposal-table slots (a power of two or 1.5 x one: any size works with hash_cell)
    int health_smem;                   // CTA-per-env variants: shared copy of the health values (see carve)
    int smem_per_env;
    unsigned char* scratch;            // BIG envs: per-env global scratch for everything but the occupancy bitmap
    long long scratch_per_env;
    double evac_reward, death_penalty, death_acc_penalty, alive_bonus;
    uint32_t* obs_wire;                // optional compact form of the observation windows (mq_env_set_obs_wire), or nullptr
};

struct DevState {
    uint32_t* pos; double* health; double* acc; uint8_t* flags; uint32_t* rmap; int* robots; int* scalars;
};

// -20.0 / (sqrt(d2) + 0.1) for d2 = 0..24: People.ROBOT_REPEL_K / (dist + 0.1), dist < ROBOT_REPEL_RANGE
// (people.py:94-95,282-284).  dist = sqrt of an exact integer, so the table is exact.
__constant__ double c_repel[25];
// map.py:11-19 MoveTO = (1,0) (0,-1) (-1,0) (0,1) (1,-1) (-1,-1) (-1,1) (1,1), packed 2 bits per direction (+1)
__device__ __forceinline__ int move_dx(int d) { return (int)((0x8246u >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int move_dy(int d) { return (int)((0xA091u >> (2 * d)) & 3u) - 1; }

// proposal-table slot (16 B): best = (priority << 32 | proposer index) min; key = target cell;
// ml = (min proposer index) | (max key+1 of the winners that left this cell) << 16
struct __align__(16) Slot { unsigned long long best; uint32_t key; uint32_t ml; };

// per-env shared memory (dist aliases the proposal table, which is dead once the moves are applied)
struct Smem {
    Slot* tab;
    double* health; double* dist; double* leaf_sum;
    uint32_t* bm; uint32_t* pos;
    int* leaf_off; int* leaf_len;
    int* nchild;         // parallel np.mean tree (wide groups): 2 * nleaf nodes in level order; leaf_sum / leaf_off / leaf_len double as
                         // the node arrays (value, offset, length)
    uint16_t* mov; uint32_t* mv;
    uint16_t* hurt;      // per-warp lists of the people standing in danger this step (T + N entries, T = threads of the group)
    uint8_t* fl;
    uint32_t* tfilt;     // target filter of the move phases: one bit per hashed target cell, 10 * nleaf words ALIASING the four leaf
                         // arrays (which the np.mean tree initialises itself, after the moves)
    int tfilt_bits;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Small envs: everything in shared memory (returns the shared bytes).  BIG envs (gbase != nullptr): only the
// occupancy bitmap and the pairwise-sum leaf tables stay in shared memory, the per-person arrays and the proposal
// table live in a per-env global scratch area (returns the shared bytes; *gbytes receives the scratch bytes).
// health_smem = false (one CTA per env, per-person arrays in shared memory): no shared copy of the health values — the chain warp
// and the guidance term read them from the env's own global rows (L2-hot: phase 1 has just touched them), which brings the CTA
// under a quarter of the SM's shared memory (4 resident CTAs instead of 3 at 1000 people).
__host__ __device__ inline size_t carve(Smem& s, unsigned char* base, unsigned char* gbase, int N, int cap, int words, int nleaf,
                                        int T, size_t* gbytes = nullptr, bool health_smem = true) {
    size_t o = 0, go = 0;
    unsigned char* pb = gbase ? gbase : base;          // where the per-person arrays go
    size_t& po = gbase ? go : o;
    const size_t table = (size_t)cap * sizeof(Slot);
    const size_t distb = sizeof(double) * (size_t)N;
    s.tab = (Slot*)(pb + po);
    s.dist = (double*)(pb + po);
    if (gbase) { po += align_up(table, 16); s.dist = (double*)(pb + po); po += align_up(distb, 16); }
    else po += align_up(table > distb ? table : distb, 16);
    s.health = nullptr;
    if (health_smem) { s.health = (double*)(pb + po); po += align_up(sizeof(double) * N, 16); }
    s.pos = (uint32_t*)(pb + po); po += align_up(sizeof(uint32_t) * N, 16);
    s.mv = (uint32_t*)(pb + po); po += align_up(sizeof(uint32_t) * N, 16);
    s.mov = (uint16_t*)(pb + po); po += align_up(sizeof(uint16_t) * N, 16);
    s.hurt = (uint16_t*)(pb + po); po += align_up(sizeof(uint16_t) * (N + T), 16);
    s.fl = (uint8_t*)(pb + po); po += align_up(N, 16);
    s.tfilt = (uint32_t*)(base + o); s.tfilt_bits = 32 * 10 * nleaf;
    s.leaf_sum = (double*)(base + o); o += sizeof(double) * 2 * nleaf;
    s.leaf_off = (int*)(base + o); o += sizeof(int) * 2 * nleaf;
    s.leaf_len = (int*)(base + o); o += sizeof(int) * 2 * nleaf;
    s.nchild = (int*)(base + o); o += sizeof(int) * 2 * nleaf;
    o = align_up(o, 16);
    s.bm = (uint32_t*)(base + o); o += sizeof(uint32_t) * align_up(words, 4);
    if (gbytes) *gbytes = align_up(go, 256);
    return align_up(o, 16);
}

// proposal-table reads: atomics act at L2, so BIG (global-memory) tables are read around L1
template <bool BIG> __device__ __forceinline__ uint32_t tab_ld(const uint32_t* p) { return BIG ? __ldcg(p) : *p; }
template <bool BIG> __device__ __forceinline__ unsigned long long tab_ld(const unsigned long long* p) { return BIG ? __ldcg(p) : *p; }

// group = the WPE warps that own one env; CW = warps per CTA.  With WPE == 8 the last warp is the "chain warp": it
// takes part in the load phases, then runs the sequential health sum while the 7 worker warps do everything else.
template <int WPE, int CW>
struct Group {
    static constexpr int SIZE = 32 * WPE;
    static constexpr bool CHAIN = WPE >= MQ_CHAIN_MIN_WPE;   // a dedicated chain warp only pays for wide groups
    static constexpr int WORKERS = CHAIN ? SIZE - 32 : SIZE;
    int gtid, gid;
    __device__ __forceinline__ Group() : gtid(threadIdx.x % SIZE), gid(threadIdx.x / SIZE) {}
    __device__ __forceinline__ void sync() const {          // all threads of the group
        if (WPE == 1) __syncwarp();
        else if (WPE == CW) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(2 * gid + 1), "r"(SIZE) : "memory");
    }
    __device__ __forceinline__ void wsync() const {         // worker threads only
        if (WPE == 1) __syncwarp();
        else if (!CHAIN && WPE == CW) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(2 * gid + 2), "r"(WORKERS) : "memory");
    }
    __device__ __forceinline__ bool worker() const { return !CHAIN || gtid < WORKERS; }
};

__device__ __forceinline__ uint32_t bm_get(const uint32_t* bm, int wpr, int x, int y) {
    return (bm[x * wpr + (y >> 5)] >> (y & 31)) & 1u;
}
__device__ __forceinline__ void bm_set(uint32_t* bm, int wpr, int x, int y) {
    atomicOr(&bm[x * wpr + (y >> 5)], 1u << (y & 31));
}
__device__ __forceinline__ void bm_clear(uint32_t* bm, int wpr, int x, int y) {
    atomicAnd(&bm[x * wpr + (y >> 5)], ~(1u << (y & 31)));
}
// slot of a cell in a table of `cap` slots: multiplicative hash, then the multiply-high range reduction (no power of two needed)
__device__ __forceinline__ uint32_t hash_cell(uint32_t c, uint32_t cap) { return __umulhi(c * 0x9E3779B1u, cap); }

// bit of a cell in the target filter (a hash independent of hash_cell's slot)
__device__ __forceinline__ uint32_t filt_bit(uint32_t c, int bits) { return __umulhi((c ^ (c >> 7)) * 0x85EBCA6Bu, (uint32_t)bits); }

__device__ __forceinline__ double box_lookup(const int* box, const double* tab, int step, int x, int y) {
    int rx = x - box[0], ry = y - box[1];
    if (rx < 0 || ry < 0 || rx >= box[2] || ry >= box[3]) return 0.0;
    return __ldg(tab + ((size_t)step * box[2] + rx) * box[3] + ry);
}

// ---------------------------------------------------------------------------------------------
// _get_state (evacuation_env.py:84-120) for every robot of the env: one lane per window CELL, six channels
// written as three 8-byte stores.  Centre of robot 0 = Map.robot_position (cx0, cy0); robots r >= 1 use
// Map.robot_positions[r] (evacuation_env_multi.py:44-53).  Executed by `nthr` threads with ids `tid`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_obs_wire(int tid, int nthr, const DevLayout& lay, const DevCfg& cfg, const Smem& sm,
                                                const int (*rob)[2], int cx0, int cy0, int fire_step, int env);
__device__ __forceinline__ void gather_obs(int tid, int nthr, const DevLayout& lay, const DevCfg& cfg, const Smem& sm,
                                           const int (*rob)[2], int cx0, int cy0, int fire_step, float* obs, double* obs64,
                                           int env) {
    const int cells = cfg.R * MQ_OBS_WIN * MQ_OBS_WIN;
    const int fs = min(fire_step, lay.n_fire_steps - 1);
    if (cfg.obs_wire) gather_obs_wire(tid, nthr, lay, cfg, sm, rob, cx0, cy0, fire_step, env);
    if (!obs && !obs64) return;
    for (int idx = tid; idx < cells; idx += nthr) {
        const int r = idx / (MQ_OBS_WIN * MQ_OBS_WIN), cell = idx - r * (MQ_OBS_WIN * MQ_OBS_WIN);
        const int i = cell / MQ_OBS_WIN, j = cell - i * MQ_OBS_WIN;
        const int cx = r == 0 ? cx0 : rob[r][0], cy = r == 0 ? cy0 : rob[r][1];
        const int mx = cx + i - 5, my = cy + j - 5;
        const bool in_grid = mx >= 0 && mx <= lay.L + 1 && my >= 0 && my <= lay.W + 1;
        const uint32_t ci = in_grid ? (uint32_t)__ldg(lay.cellinfo + mx * lay.stride + my) : 2u;   // off-grid: blocked
        const double v1 = (ci & 1u) ? (double)bm_get(sm.bm, lay.wpr, mx, my) : 0.0;      // rmap if Check_Valid
        const double v2 = box_lookup(lay.int_box, lay.danger_int, fs, mx, my);           // danger at integer coords
        const double v3 = (ci & 2u) ? 1.0 : 0.0, v4 = (ci & 4u) ? 1.0 : 0.0;
        const double v5 = (i == 5 && j == 5) ? 1.0 : 0.0;
        const size_t o = ((size_t)env * cfg.R + r) * MQ_OBS_SIZE + (size_t)cell * MQ_OBS_CH;
        if (obs) {                                    // = state.astype(np.float32) at dqn_agent.py:109
            float2* dst = reinterpret_cast<float2*>(obs + o);
            dst[0] = make_float2(0.f, (float)v1);     // channel 0 is space/inf == 0 (quirk Q1)
            dst[1] = make_float2((float)v2, (float)v3);
            dst[2] = make_float2((float)v4, (float)v5);
        }
        if (obs64) {
            double* d = obs64 + o;
            d[0] = 0.0; d[1] = v1; d[2] = v2; d[3] = v3; d[4] = v4; d[5] = v5;
        }
    }
}

// Compact wire form of the observation windows for the host-buffer interface (mq_env_set_obs_wire): MQ_OBS_WIRE_WORDS = 136
// words (544 B) per (env, robot) window instead of 2904 B.  Words 0..120 = channel 2 (danger, f32 bits) of cell c = i*11+j;
// 121..124 / 125..128 / 129..132 = channels 1 / 3 / 4 as bit planes (bit c); 133..135 = 0.  Channel 0 is identically 0
// (quirk Q1) and channel 5 marks the centre cell (evacuation_env.py:116-117): neither travels.  One lane per cell of a window
// padded to 128 cells, so a warp's ballots ARE the plane words.  nthr and tid's warp must be whole warps.
__device__ __forceinline__ void gather_obs_wire(int tid, int nthr, const DevLayout& lay, const DevCfg& cfg, const Smem& sm,
                                                const int (*rob)[2], int cx0, int cy0, int fire_step, int env) {
    const int fs = min(fire_step, lay.n_fire_steps - 1);
    for (int idx = tid; idx < cfg.R * 128; idx += nthr) {
        const int r = idx >> 7, cell = idx & 127;
        const bool in_win = cell < MQ_OBS_WIN * MQ_OBS_WIN;
        uint32_t ci = 0u, b1 = 0u;
        float v2 = 0.f;
        if (in_win) {
            const int i = cell / MQ_OBS_WIN, j = cell - i * MQ_OBS_WIN;
            const int cx = r == 0 ? cx0 : rob[r][0], cy = r == 0 ? cy0 : rob[r][1];
            const int mx = cx + i - 5, my = cy + j - 5;
            const bool in_grid = mx >= 0 && mx <= lay.L + 1 && my >= 0 && my <= lay.W + 1;
            ci = in_grid ? (uint32_t)__ldg(lay.cellinfo + mx * lay.stride + my) : 2u;
            b1 = (ci & 1u) ? bm_get(sm.bm, lay.wpr, mx, my) : 0u;
            v2 = (float)box_lookup(lay.int_box, lay.danger_int, fs, mx, my);
        }
        const uint32_t p1 = __ballot_sync(0xFFFFFFFFu, b1 != 0u);
        const uint32_t p3 = __ballot_sync(0xFFFFFFFFu, (ci & 2u) != 0u);
        const uint32_t p4 = __ballot_sync(0xFFFFFFFFu, (ci & 4u) != 0u);
        uint32_t* dst = cfg.obs_wire + ((size_t)env * cfg.R + r) * MQ_OBS_WIRE_WORDS;
        if (in_win) dst[cell] = __float_as_uint(v2);
        else if (cell >= 125) dst[133 + (cell - 125)] = 0u;
        if ((cell & 31) == 0) { const int w = cell >> 5; dst[121 + w] = p1; dst[125 + w] = p3; dst[129 + w] = p4; }
    }
}

__device__ __forceinline__ void store_bitmap(int tid, int nthr, const DevLayout& lay, const Smem& sm, uint32_t* g_rmap, int env) {
    uint4* dst = reinterpret_cast<uint4*>(g_rmap + (size_t)env * lay.rmap_words);
    const uint4* src = reinterpret_cast<const uint4*>(sm.bm);
    for (int w = tid; w < lay.rmap_words / 4; w += nthr) dst[w] = src[w];
}

// ---------------------------------------------------------------------------------------------
// EvacuationEnv.reset (evacuation_env.py:61-82) + People.__init__ spawn (people.py:185-194), group-wide.
// sc = shared copy of the env scalars, rob = shared copy of robot_positions.
// ---------------------------------------------------------------------------------------------
template <int WPE, int CW>
__device__ void reset_env(const Group<WPE, CW>& g, const DevLayout& lay, const DevCfg& cfg, const DevState& st, const Smem& sm,
                          int* sc, int (*rob)[2], const int16_t* inject, float* obs, double* obs64, int env) {
    constexpr int T = Group<WPE, CW>::SIZE;
    const int N = cfg.N, tid = g.gtid;
    for (int w = tid; w < lay.rmap_words; w += T) sm.bm[w] = 0u;
    g.sync();
    const uint32_t env_id = (uint32_t)(cfg.env_id_base + env);
    const uint32_t episode = (uint32_t)sc[MQ_S_EPISODE];
    const size_t base = (size_t)env * cfg.n_pad;
    for (int i = tid; i < N; i += T) {
        int x, y;
        if (inject) {
            x = inject[((size_t)env * N + i) * 2];
            y = inject[((size_t)env * N + i) * 2 + 1];
        } else {
            for (uint32_t attempt = 0;; ++attempt) {   // randint(1, L-2), randint(1, W-2) until Check_Valid
                uint4 w = philox4x32(env_id, episode, (uint32_t)i, STREAM_SPAWN + (attempt >> 1), cfg.pk);
                uint32_t wx = (attempt & 1) ? w.z : w.x, wy = (attempt & 1) ? w.w : w.y;
                x = 1 + (int)__umulhi(wx, (uint32_t)(lay.L - 2));
                y = 1 + (int)__umulhi(wy, (uint32_t)(lay.W - 2));
                if (__ldg(lay.cellinfo + x * lay.stride + y) & 1u) break;
            }
        }
        st.pos[base + i] = (uint32_t)x | ((uint32_t)y << 16);
        st.health[base + i] = 100.0;      // people.py:19
        st.acc[base + i] = 0.0;           // people.py:22
        st.flags[base + i] = 0;
        bm_set(sm.bm, lay.wpr, x, y);     // rmap[x][y] = 1, duplicates allowed (quirk Q2)
    }
    g.sync();
    if (tid == 0) {
        if (cfg.reset_robots) {           // evacuation_env_multi.py:35-36
            for (int r = 0; r < cfg.R; ++r) { rob[r][0] = lay.robot_start[r][0]; rob[r][1] = lay.robot_start[r][1]; }
            sc[MQ_S_ROBOT_POS_X] = rob[0][0]; sc[MQ_S_ROBOT_POS_Y] = rob[0][1];
        } else {                          // evacuation_env.py:64 — robot_positions untouched (quirk Q7)
            sc[MQ_S_ROBOT_POS_X] = lay.reset_center[0]; sc[MQ_S_ROBOT_POS_Y] = lay.reset_center[1];
        }
        if (cfg.reset_fire) sc[MQ_S_FIRE_STEP] = 0;
        sc[MQ_S_CUR_STEP] = 0; sc[MQ_S_PREV_EVAC] = 0; sc[MQ_S_PREV_DEAD] = 0;
        sc[MQ_S_EVAC] = 0; sc[MQ_S_DEAD] = 0;
        sc[MQ_S_EPISODE] = (int)(episode + 1);
    }
    g.sync();
    if (tid < MQ_ENV_SCALARS) st.scalars[(size_t)env * MQ_ENV_SCALARS + tid] = sc[tid];
    if (tid < MAXR * 2) st.robots[(size_t)env * MAXR * 2 + tid] = rob[tid >> 1][tid & 1];
    store_bitmap(tid, T, lay, sm, st.rmap, env);
    gather_obs(tid, T, lay, cfg, sm, rob, sc[MQ_S_ROBOT_POS_X], sc[MQ_S_ROBOT_POS_Y], sc[MQ_S_FIRE_STEP], obs, obs64, env);
}

template <int WPE, int CW, bool BIG, bool MULTI>
__global__ void __launch_bounds__(32 * CW)
env_reset_kernel(DevLayout lay_in, DevCfg cfg, DevState st, const uint8_t* env_mask, const int16_t* inject, float* obs,
                 double* obs64) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GROUPS = CW / WPE;
    __shared__ int s_sc[GROUPS][MQ_ENV_SCALARS];
    __shared__ int s_rob[GROUPS][MAXR][2];
    const Group<WPE, CW> g;
    const int env = blockIdx.x * GROUPS + g.gid;
    if (env >= cfg.n_envs) return;
    if (env_mask && !env_mask[env]) return;
    DevLayout lay = lay_in;
    select_layout<MULTI>(lay, env);
    Smem sm;
    carve(sm, smem_raw + (size_t)g.gid * cfg.smem_per_env, BIG ? cfg.scratch + (size_t)env * cfg.scratch_per_env : nullptr, cfg.N,
          cfg.hash_cap, lay.rmap_words, cfg.n_leaf_max, Group<WPE, CW>::SIZE, nullptr, !(WPE >= 8 && !BIG) || cfg.health_smem);
    int* sc = s_sc[g.gid];
    int(*rob)[2] = s_rob[g.gid];
    if (g.gtid < MQ_ENV_SCALARS) sc[g.gtid] = st.scalars[(size_t)env * MQ_ENV_SCALARS + g.gtid];
    if (g.gtid < MAXR * 2) rob[g.gtid >> 1][g.gtid & 1] = st.robots[(size_t)env * MAXR * 2 + g.gtid];
    g.sync();
    reset_env<WPE, CW>(g, lay, cfg, st, sm, sc, rob, inject, obs, obs64, env);
}

// numpy pairwise summation tree (np.mean at evacuation_env.py:228): leaves are blocks of <= 128 elements
// summed with 8 interleaved accumulators, inner nodes split at n/2 rounded down to a multiple of 8.
__device__ int enumerate_leaves(int n, int* off, int* len) {
    if (n <= 128) { off[0] = 0; len[0] = n; return 1; }
    int so[24], sn[24], sp = 0, nl = 0;
    so[0] = 0; sn[0] = n; sp = 1;
    while (sp) {
        --sp;
        int o = so[sp], m = sn[sp];
        if (m <= 128) { off[nl] = o; len[nl] = m; ++nl; }
        else {
            int n2 = m / 2; n2 -= n2 % 8;
            so[sp] = o + n2; sn[sp] = m - n2; ++sp;     // right, popped after
            so[sp] = o; sn[sp] = n2; ++sp;              // left, popped first
        }
    }
    return nl;
}
__device__ double combine_leaves(int n, const double* leaf_sum) {
    if (n <= 128) return leaf_sum[0];
    int sn[40]; signed char sk[40]; double val[24];
    int sp = 0, vp = 0, next = 0;
    sn[0] = n; sk[0] = 0; sp = 1;
    while (sp) {
        --sp;
        int m = sn[sp];
        if (sk[sp]) { double b = val[--vp]; double a = val[--vp]; val[vp++] = a + b; }
        else if (m <= 128) val[vp++] = leaf_sum[next++];
        else {
            int n2 = m / 2; n2 -= n2 % 8;
            sk[sp] = 1; ++sp;
            sn[sp] = m - n2; sk[sp] = 0; ++sp;
            sn[sp] = n2; sk[sp] = 0; ++sp;
        }
    }
    return val[0];
}
// one leaf by a group of 8 lanes (lane g = accumulator r[g]); result valid in group lane 0
__device__ __forceinline__ double leaf_sum8(const double* a, int n, int g, uint32_t gmask) {
    if (n < 8) {
        double res = 0.;
        if (g == 0) for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r = a[g];
    int main_end = n - (n % 8);
    for (int i = 8; i < main_end; i += 8) r += a[i + g];
    double r1 = r + __shfl_down_sync(gmask, r, 1, 8);       // lanes 0,2,4,6: r[g]+r[g+1]
    double r2 = r1 + __shfl_down_sync(gmask, r1, 2, 8);     // lanes 0,4: (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
    double res = r2 + __shfl_down_sync(gmask, r2, 4, 8);    // lane 0
    if (g == 0) for (int i = main_end; i < n; ++i) res += a[i];
    return res;
}
// sum(p.health for p in list if not p.dead) (evacuation_env.py:245): strictly left to right; dead entries hold +0.0.
// A literal chain costs ~20 cycles per person (dependent fp64 adds).  Most people are unhurt, i.e. exactly 100.0, and a run
// of k additions of 100.0 to S is EXACT as long as S stays inside its binade (100 k is a multiple of S's ulp for S < 2^54):
// the rounded left-to-right result of such a run is S + 100 j for the j additions that stay below the next power of two,
// one real (rounding) addition across it, and so on — O(binades) instead of O(k).  Hurt people are added one by one.
__device__ __forceinline__ double add_hundreds(double S, int k) {
    while (k > 0) {
        if (S == 0.0) { S = 100.0; --k; continue; }
        const int e = (__double2hiint(S) >> 20) & 0x7FF;                // biased exponent (S > 0, finite)
        const double P = __hiloint2double((e + 1) << 20, 0);             // the next power of two above S
        const double d = P - S;                                          // exact
        long long j = (long long)(d / 100.0);
        while ((double)j * 100.0 >= d) --j;
        while ((double)(j + 1) * 100.0 < d) ++j;                         // now 100 j < P - S <= 100 (j + 1)
        if (j >= k) { S = S + (double)k * 100.0; k = 0; }
        else { S = S + (double)j * 100.0; S = S + 100.0; k -= (int)j + 1; }
    }
    return S;
}
// the literal chain (one lane): cheapest for the small warp-per-env shapes, where crossing ~8 binades costs more than 150 adds
__device__ __forceinline__ double health_chain_literal(const double* h, int N) {
    double tot = 0.0;
    int i = 0;
    for (; i + 4 <= N; i += 4) {
        const double2 a = *reinterpret_cast<const double2*>(h + i);
        const double2 b = *reinterpret_cast<const double2*>(h + i + 2);
        tot += a.x; tot += a.y; tot += b.x; tot += b.y;
    }
    for (; i < N; ++i) tot += h[i];
    return tot;
}
// run-based chain, executed by a whole warp; every lane returns the sum
// FLAGS: h is the env's global health row and dead people (flags bit 1, read from fl) count as + 0.0; otherwise h already holds
// the summands.
template <int DEPTH, bool FLAGS = false>      // batches of 32 fetched before any is consumed: 8 when the list lives in global memory, else 1
__device__ __noinline__ double health_chain_runs(const double* h, int N, int lane, const uint8_t* fl = nullptr) {
    double S = 0.0;
    int pend = 0;                                   // 100.0's seen since the last hurt person
    for (int b0 = 0; b0 < N; b0 += 32 * DEPTH) {
        double vv[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            const int i = b0 + 32 * u + lane;
            if (FLAGS) vv[u] = (i < N && !(fl[i] & 2u)) ? __ldcg(h + i) : 0.0;
            else vv[u] = i < N ? h[i] : 0.0;
        }
        uint32_t m100_[DEPTH], hard_[DEPTH], any_hard = 0;
        int n100 = 0;
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {           // classify all fetched batches first (independent ballots)
            m100_[u] = __ballot_sync(0xFFFFFFFFu, vv[u] == 100.0);
            hard_[u] = __ballot_sync(0xFFFFFFFFu, vv[u] != 100.0 && vv[u] != 0.0);      // + 0.0 (dead / past the end) is a no-op
            any_hard |= hard_[u];
            n100 += __popc(m100_[u]);
        }
        if (!any_hard) { pend += n100; continue; }  // nobody hurt in these 32 * DEPTH people
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            const int b = b0 + 32 * u;
            if (b >= N) break;
            const double v = vv[u];
            const uint32_t m100 = m100_[u];
            uint32_t hard = hard_[u];
            if (__popc(hard) > 4) {                     // mostly hurt people: the plain chain over the batch is cheaper
                S = add_hundreds(S, pend);
                pend = 0;
                const int cnt = min(32, N - b);
                for (int k = 0; k < cnt; ++k) S = S + __shfl_sync(0xFFFFFFFFu, v, k);
                continue;
            }
            uint32_t done = 0;
            while (hard) {
                const int hp = __ffs(hard) - 1;
                const uint32_t below = (1u << hp) - 1u;
                S = add_hundreds(S, pend + __popc(m100 & below & ~done));
                pend = 0;
                S = S + __shfl_sync(0xFFFFFFFFu, v, hp);
                done |= below | (1u << hp);
                hard &= hard - 1;
            }
            pend += __popc(m100 & ~done);
        }
    }
    return add_hundreds(S, pend);
}

// ---------------------------------------------------------------------------------------------
// The fused step.
// ---------------------------------------------------------------------------------------------
#ifndef MQ_PF_MAX
#define MQ_PF_MAX 5
#endif
#ifndef MQ_PF_WIDE
#define MQ_PF_WIDE 2      // persons per thread prefetched in phase 1 of the CTA-per-env variants (4: 0.752 ms at C3, 2: 0.718, 1: 0.722 — registers)
#endif
constexpr int PF_MAX = MQ_PF_MAX;  // persons per thread whose state is fetched before any of them is processed (5 x 32 >= 150: one pass at C2)
#ifdef MQ_ENV_TRACE
__device__ long long g_env_trace[16];
#define ENV_MARK(k) do { if (tid == 0) { const long long _t = clock64(); atomicAdd((unsigned long long*)&g_env_trace[k], (unsigned long long)(_t - _tprev)); _tprev = _t; } } while (0)
#else
#define ENV_MARK(k) do { } while (0)
#endif
#ifndef MQ_SCORE_U_SMALL
#define MQ_SCORE_U_SMALL 2
#endif
constexpr int SCORE_UNROLL = MQ_SCORE_U_SMALL;     // movers scored per lane and iteration of phase 2 (warp-per-env and BIG variants)
// build-time tuning knobs of the CTA-per-env variant (dqn_marl_b200/build.py build_variant): resident CTAs per SM the register
// allocation aims at, and the scoring unroll
#ifndef MQ_CTA8_PER_SM
#define MQ_CTA8_PER_SM 4
#endif
#ifndef MQ_SCORE_U8
#define MQ_SCORE_U8 1      // CTA-per-env variant: with the lazy noise one mover per lane is fastest (fewer live registers: 0.759 vs 0.794 ms at C3)
#endif

template <int WPE, int CW, bool BIG, bool MULTI>
__global__ void __launch_bounds__(32 * CW, WPE < 8 ? 28 / CW : (BIG ? 1 : MQ_CTA8_PER_SM))      // 28 env-warps / 4 env-CTAs resident per SM (BIG: one)
env_step_kernel(DevLayout lay_in, DevCfg cfg, DevState st, const int* __restrict__ actions, float* obs, double* obs64,
                double* reward_out, uint8_t* done_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using G = Group<WPE, CW>;
    constexpr int GROUPS = CW / WPE;
    constexpr int T = G::SIZE;          // threads of the group
    constexpr int TW = G::WORKERS;      // threads that do the parallel phases after phase 1
    __shared__ int s_sc[GROUPS][MQ_ENV_SCALARS];
    __shared__ int s_rob[GROUPS][MAXR][2];
    __shared__ int s_cnt_all[GROUPS][6];        // evacuated, dead, guidance in halves, live, movers, reset flag
    __shared__ int s_wtot_all[GROUPS][WPE];
    __shared__ double s_sum_all[GROUPS][2];     // sum of distances, total health
    __shared__ int s_pre[GROUPS + 1];           // COOP: exclusive prefix of the envs' scoring chunks
    __shared__ int s_lvl_all[WPE >= 8 ? GROUPS : 1][20];   // wide groups: first node of every level of the np.mean tree, [19] = levels
    __shared__ const double* s_dp5[MULTI ? GROUPS : 1];    // MULTI + COOP: the dp5 table of every env of the CTA

    const G g;
    const int env = blockIdx.x * GROUPS + g.gid;
    if (env >= cfg.n_envs) return;              // whole group leaves together (GROUPS == 1 when WPE == CW)
    DevLayout lay = lay_in;
    select_layout<MULTI>(lay, env);
    if (MULTI && g.gtid == 0) s_dp5[g.gid] = lay.dp5;
#ifdef MQ_ENV_TRACE
    long long _tprev = clock64();
#endif
    const bool HSM = !(WPE >= 8 && !BIG) || cfg.health_smem;     // shared copy of the health values (see carve)
    Smem sm;
    carve(sm, smem_raw + (size_t)g.gid * cfg.smem_per_env, BIG ? cfg.scratch + (size_t)env * cfg.scratch_per_env : nullptr, cfg.N,
          cfg.hash_cap, lay.rmap_words, cfg.n_leaf_max, Group<WPE, CW>::SIZE, nullptr, HSM);
    int* sc = s_sc[g.gid];
    int(*rob)[2] = s_rob[g.gid];
    int* s_cnt = s_cnt_all[g.gid];
    int* s_wtot = s_wtot_all[g.gid];
    double* s_sum = s_sum_all[g.gid];

    const int tid = g.gtid;
    const int lane = tid & 31, warp = tid >> 5;
    const int N = cfg.N, stride = lay.stride, wpr = lay.wpr;
    const size_t base = (size_t)env * cfg.n_pad;
    const uint32_t hcap = (uint32_t)cfg.hash_cap;

    // ---- stage: scalars, robots, occupancy bitmap; clear the proposal table ------------------------
    // Nothing here makes the group WAIT for global memory: the bitmap comes in by asynchronous copies that are only awaited at
    // the barrier after phase 1 (its first readers are the scoring lanes), the scalars and robots are fetched by warp 0 for the
    // later phases while every thread reads the two scalars phase 1 needs (tick, fire step) itself, together with its person
    // state.  The staging round trip (8 % of an env-step, clock64 trace) now overlaps phase 1.
    if (tid < 6) s_cnt[tid] = 0;
    g.sync();                                       // only the zeroed counters are behind this barrier
    {
        const uint4* src = reinterpret_cast<const uint4*>(st.rmap + (size_t)env * lay.rmap_words);
        uint4* dst = reinterpret_cast<uint4*>(sm.bm);
        for (int w = tid; w < lay.rmap_words / 4; w += T)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + w)), "l"(src + w) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint4 empty = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, HEMPTY, 0x0000FFFFu);   // best = ~0, key, ml = (min 0xFFFF, leave 0)
        uint4* tab = reinterpret_cast<uint4*>(sm.tab);
        for (int h = tid; h < cfg.hash_cap; h += T) tab[h] = empty;
        for (int w = tid; w < sm.tfilt_bits / 32; w += T) sm.tfilt[w] = 0u;
    }
    const uint32_t env_id = (uint32_t)(cfg.env_id_base + env);
    const uint32_t tick = (uint32_t)st.scalars[(size_t)env * MQ_ENV_SCALARS + MQ_S_TICK];
    const int fire_step = min(st.scalars[(size_t)env * MQ_ENV_SCALARS + MQ_S_FIRE_STEP], lay.n_fire_steps - 1);
    if (warp == 0) {
        if (lane < MQ_ENV_SCALARS) sc[lane] = st.scalars[(size_t)env * MQ_ENV_SCALARS + lane];
        if (lane < MAXR * 2) rob[lane >> 1][lane & 1] = st.robots[(size_t)env * MAXR * 2 + lane];
        __syncwarp();
        // ---- Map.move_robot for every robot, in robot order (map.py:160-202; evacuation_env_multi.py:60-63).
        //      Nobody reads the robots or the staged scalars before the barrier that follows phase 1. ----
        if (lane == 0) {
            for (int r = 0; r < cfg.R; ++r) {
                int a = actions[(size_t)env * cfg.R + r];
                if (a < 0 || a > 4) continue;                       // map.py:180-181 (returns before the re-alias)
                int x = rob[r][0], y = rob[r][1], nx = x, ny = y;
                if (a == 0) nx = x + 1; else if (a == 1) ny = y - 1; else if (a == 2) nx = x - 1; else if (a == 3) ny = y + 1;
                bool ok = lay.robot_range[0] <= nx && nx <= lay.robot_range[1] && 0 <= ny && ny <= lay.W;
                // Check_Valid (map.py:85-92): inside 1..L x 1..W and finite potential
                ok = ok && nx >= 1 && nx <= lay.L && ny >= 1 && ny <= lay.W && (__ldg(lay.cellinfo + nx * stride + ny) & 1u);
                if (ok) { rob[r][0] = nx; rob[r][1] = ny; }
                if (r == 0) { sc[MQ_S_ROBOT_POS_X] = rob[0][0]; sc[MQ_S_ROBOT_POS_Y] = rob[0][1]; }   // map.py:200-201
            }
        }
    }
    ENV_MARK(0);      // stage

    // ---- phase 1 (people.py:203-220): health, speed, accumulator; movers are compacted.
    //      State of PF persons per thread is fetched up front so the DRAM round trips overlap.  The few people that stand
    //      in danger are only LISTED here (per warp) and get their health update in a compact second pass: the keyed draw
    //      and the loss arithmetic then run once per 32 of them instead of once per loop iteration. ----
    uint16_t* const my_hurt = sm.hurt + (size_t)warp * (((N + T - 1) / T) * 32);
    int n_hurt = 0;                              // warp-uniform
    // Person.update_state speed (people.py:38-44), accumulator (people.py:216-220), shared-memory copies; returns "moves"
    auto advance = [&](int i, uint32_t fl, double h, double a) -> bool {
        bool mover = false;
        if (!(fl & 2u)) {
            double hr = 1.0;                              // 100.0 / 100.0; most people are unhurt: skip the fp64 division
            if (h != 100.0) hr = h / 100.0;
            const double speed = (h < 20.0) ? 0.4 : 1.0 * (0.3 + 0.7 * hr);
            a += speed * 0.5;
            if (a >= 1.0) { a -= 1.0; mover = true; }
            st.acc[base + i] = a;
        }
        sm.fl[i] = (uint8_t)fl;
        if (HSM) sm.health[i] = (fl & 2u) ? 0.0 : h;          // summand of evacuation_env.py:245 (dead -> +0.0)
        return mover;
    };
    auto push_movers = [&](bool mover, int i) {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mover);
        if (bal) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&s_cnt[4], __popc(bal));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            if (mover) {
                const int mi = wbase + __popc(bal & ((1u << lane) - 1u));
                sm.mov[mi] = (uint16_t)i;
                sm.mv[mi] = 0xFFFFFFFFu;
            }
        }
    };
    constexpr int PF = WPE == 1 ? PF_MAX : MQ_PF_WIDE;
    for (int i0 = 0; i0 < N; i0 += T * PF) {
        uint32_t p_[PF], fl_[PF];
        double h_[PF], a_[PF], dg_[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = i0 + k * T + tid;
            p_[k] = 0; fl_[k] = 3u; h_[k] = 0.0; a_[k] = 0.0;
            if (i < N) { p_[k] = st.pos[base + i]; fl_[k] = st.flags[base + i]; h_[k] = st.health[base + i]; a_[k] = st.acc[base + i]; }
        }
#pragma unroll
        for (int k = 0; k < PF; ++k)
            dg_[k] = (fl_[k] & 3u) ? 0.0 : box_lookup(lay.ctr_box, lay.danger_ctr, fire_step, (int)(p_[k] & 0xFFFFu), (int)(p_[k] >> 16));
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = i0 + k * T + tid;
            if (i0 + k * T >= N) break;                       // uniform over the warp
            bool mover = false, hurt = false;
            if (i < N) {
                const uint32_t fl = fl_[k];
                sm.pos[i] = p_[k];
                if (fl & 3u) { sm.fl[i] = (uint8_t)fl; if (HSM) sm.health[i] = (fl & 2u) ? 0.0 : h_[k]; }
                else if (dg_[k] > 0.0) hurt = true;
                else mover = advance(i, fl, h_[k], a_[k]);
            }
            const uint32_t hb = __ballot_sync(0xFFFFFFFFu, hurt);
            if (hb) {
                if (hurt) my_hurt[n_hurt + __popc(hb & ((1u << lane) - 1u))] = (uint16_t)i;
                n_hurt += __popc(hb);
            }
            push_movers(mover, i);
        }
    }
    __syncwarp();
    for (int e0 = 0; e0 < n_hurt; e0 += 32) {                 // Person.update_health (people.py:61-88), compact
        const int e = e0 + lane;
        bool mover = false;
        int i = 0;
        if (e < n_hurt) {
            i = my_hurt[e];
            const uint32_t p = sm.pos[i];
            double h = st.health[base + i];
            const double a = st.acc[base + i];
            uint32_t fl = 0u;
            const double danger = box_lookup(lay.ctr_box, lay.danger_ctr, fire_step, (int)(p & 0xFFFFu), (int)(p >> 16));
            const uint4 w4 = philox4x32(env_id, tick, (uint32_t)i, STREAM_HEALTH, cfg.pk);
            const double u = u53(w4.x, w4.y);
            double loss;
            if (danger >= 0.8) loss = danger * 50.0 + (1.0 + (3.0 - 1.0) * u);
            else if (danger >= 0.5) loss = danger * 40.0 + (0.8 + (2.0 - 0.8) * u);
            else if (danger >= 0.2) loss = danger * 30.0 + (0.5 + (1.5 - 0.5) * u);
            else loss = danger * 20.0 + (0.2 + (1.0 - 0.2) * u);
            if (h < 50.0) loss *= 1.2;
            h -= loss;
            if (h <= 0.0) { h = 0.0; fl |= 2u; } else if (h <= 8.0) fl |= 2u;
            h = fmax(0.0, fmin(h, 100.0));
            st.health[base + i] = h;
            if (fl & 2u) st.flags[base + i] = (uint8_t)fl;
            mover = advance(i, fl, h, a);
        }
        push_movers(mover, i);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");           // this thread's pieces of the occupancy bitmap have landed
    g.sync();

    ENV_MARK(1);      // phase 1
    if (G::CHAIN && !g.worker()) {
        // chain warp: the left-to-right health sum only needs phase 1; it overlaps with everything the workers do
        const double th = HSM ? health_chain_runs<BIG ? 8 : 1>(sm.health, N, lane)
                              : health_chain_runs<4, true>(st.health + base, N, lane, sm.fl);
        if (lane == 0) s_sum[1] = th;
    } else {
        const int wt = tid;                 // worker thread id (workers are the first TW threads of the group)
        const int n_mov = s_cnt[4];

        // ---- phase 2 (people.py:221-230, 255-297): 4 lanes score the 8 directions of one mover ----------------
        // Warp-per-env CTAs with many envs (COOP): the cost of this phase differs by 10x between envs (everybody moves /
        // nobody moves) and the launch is one wave, so the slowest env of the whole grid sets the kernel time.  The warps of
        // a CTA therefore score the movers of ALL its envs together: the envs' 32-item chunks are numbered through an
        // exclusive prefix and dealt to the warps round-robin, between two CTA barriers.
        constexpr bool COOP = WPE == 1 && CW >= 8;
        const bool coop = COOP && (blockIdx.x + 1) * GROUPS <= cfg.n_envs;       // CTA-uniform; a partial last CTA works per warp
        if (coop) {
            __syncthreads();                       // every env of the CTA has its mover list and its robots
            if (g.gid == 0) {
                const int nch = lane < GROUPS ? (s_cnt_all[lane][4] * 4 + 31) >> 5 : 0;
                int incl = nch;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
                if (lane < GROUPS) s_pre[lane] = incl - nch;
                if (lane == GROUPS - 1) s_pre[GROUPS] = incl;
            }
            __syncthreads();
            const int total = s_pre[GROUPS];
            const int q = lane & 3, R = cfg.R;
            constexpr int U = SCORE_UNROLL;
            for (int c0 = g.gid; c0 < total; c0 += U * GROUPS) {
                int mi[U], i[U], x[U], y[U], rx0[U], ry0[U], eo[U];
                bool act[U];
                double dpv[U][2];
                const uint16_t* movp[U]; const uint32_t* posp[U]; const uint32_t* bmp[U]; uint32_t* mvp[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u * GROUPS;
                    int e = 0;
                    if (c < total) {                // owner env of chunk c: last e with s_pre[e] <= c (<= 32 entries, warp-uniform)
                        int lo = 0, hi = GROUPS - 1;
                        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pre[mid] <= c) lo = mid; else hi = mid - 1; }
                        e = lo;
                    }
                    eo[u] = e;
                    const long long delta = (long long)(e - g.gid) * cfg.smem_per_env;
                    movp[u] = reinterpret_cast<const uint16_t*>(reinterpret_cast<const unsigned char*>(sm.mov) + delta);
                    posp[u] = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(sm.pos) + delta);
                    bmp[u] = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(sm.bm) + delta);
                    mvp[u] = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(sm.mv) + delta);
                    rx0[u] = s_rob[e][0][0]; ry0[u] = s_rob[e][0][1];
                    const int it = (c - s_pre[e]) * 32 + lane;
                    mi[u] = it >> 2;
                    act[u] = c < total && mi[u] < s_cnt_all[e][4];
                    i[u] = act[u] ? (int)movp[u][mi[u]] : 0;
                    const uint32_t p = act[u] ? posp[u][i[u]] : 0x00010001u;
                    x[u] = (int)(p & 0xFFFFu); y[u] = (int)(p >> 16);
                    dpv[u][0] = -INFINITY; dpv[u][1] = -INFINITY;
                    if (act[u]) {
                        const double* dp5e = MULTI ? s_dp5[e] : lay.dp5;
                        const double2 v = __ldg(reinterpret_cast<const double2*>(dp5e + (size_t)(x[u] * stride + y[u]) * 8) + q);
                        dpv[u][0] = v.x; dpv[u][1] = v.y;
                    }
                }
                // deterministic scores first; the keyed noise is drawn only when some mover of the chunk has two or more
                // directions within 0.2 of its best one (see LAZY NOISE below); warp-uniform decision per chunk
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bool adm[2];
                    double det[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int d = q * 2 + k;
                        const int nx = x[u] + move_dx(d), ny = y[u] + move_dy(d);
                        adm[k] = (dpv[u][k] > -INFINITY) && !bm_get(bmp[u], wpr, nx, ny);
                        int d2;
                        {
                            const int ax = min(abs(nx - rx0[u]), 30000), ay = min(abs(ny - ry0[u]), 30000);
                            d2 = ax * ax + ay * ay;
                        }
                        if (R > 1) {                         // warp-uniform
                            for (int r = 1; r < R; ++r) {
                                const int ax = min(abs(nx - s_rob[eo[u]][r][0]), 30000), ay = min(abs(ny - s_rob[eo[u]][r][1]), 30000);
                                d2 = min(d2, ax * ax + ay * ay);
                            }
                        }
                        const double eff = d2 < 25 ? c_repel[d2] : 0.0;
                        det[k] = dpv[u][k] + eff;
                    }
                    double gmax = fmax(adm[0] ? det[0] : -INFINITY, adm[1] ? det[1] : -INFINITY);
                    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 1));
                    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 2));
                    const bool c0 = adm[0] && det[0] >= gmax - 0.25, c1 = adm[1] && det[1] >= gmax - 0.25;
                    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, c0), b1 = __ballot_sync(0xFFFFFFFFu, c1);
                    const int sh = lane & 28;
                    const int cnt = __popc((b0 >> sh) & 0xFu) + __popc((b1 >> sh) & 0xFu);
                    if (!__any_sync(0xFFFFFFFFu, act[u] && cnt >= 2)) {
                        if (act[u] && cnt == 1 && (c0 || c1)) mvp[u][mi[u]] = (uint32_t)(q * 2 + (c1 ? 1 : 0)) << 20;
                        continue;
                    }
                    const uint4 w = philox4x32((uint32_t)(cfg.env_id_base + blockIdx.x * GROUPS + eo[u]), (uint32_t)s_sc[eo[u]][MQ_S_TICK],
                                               (uint32_t)i[u], (uint32_t)q, cfg.pk);
                    double best_score = -INFINITY;
                    int best_dir = 8;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double un = k ? u53(w.z, w.w) : u53(w.x, w.y);
                        const double noise = -0.1 + (0.1 - -0.1) * un;       // random.uniform(-0.1, 0.1)
                        const double score = det[k] + noise;                   // people.py:287-291
                        if (adm[k] && score > best_score) { best_score = score; best_dir = q * 2 + k; }
                    }
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        const double os = __shfl_xor_sync(0xFFFFFFFFu, best_score, o);
                        const int od = __shfl_xor_sync(0xFFFFFFFFu, best_dir, o);
                        if (os > best_score || (os == best_score && od < best_dir)) { best_score = os; best_dir = od; }
                    }
                    if (act[u] && q == 0 && best_dir < 8) mvp[u][mi[u]] = (uint32_t)best_dir << 20;
                }
            }
            __syncthreads();
        }
        if (!coop) {
            const int rbx0 = rob[0][0], rby0 = rob[0][1];        // robot 0 in registers; further robots (R > 1) are read from shared memory
            const int q = lane & 3, R = cfg.R;
            const int n_items = (n_mov * 4 + 31) & ~31;          // whole warps take part in the shuffles
            // admissibility (Check_Valid through dp5 = -inf, rmap == 0) and the deterministic part of the score of direction
            // d = 2q + k: delta_p * 5.0 + robot_effect (people.py:268-288)
            auto det_score = [&](int xx, int yy, double dp, int d, bool& adm) -> double {
                const int nx = xx + move_dx(d), ny = yy + move_dy(d);
                adm = (dp > -INFINITY) && !bm_get(sm.bm, wpr, nx, ny);
                // squared distance to the nearest robot, coordinates saturated at 30000 (robots may sit far off-map:
                // evaluate_strategies.py:83 sets [1000,1000]; anything >= 25 means "out of range")
                int d2;
                {
                    const int ax = min(abs(nx - rbx0), 30000), ay = min(abs(ny - rby0), 30000);
                    d2 = ax * ax + ay * ay;
                }
                if (R > 1) {                         // warp-uniform
                    for (int r = 1; r < R; ++r) {
                        const int ax = min(abs(nx - rob[r][0]), 30000), ay = min(abs(ny - rob[r][1]), 30000);
                        d2 = min(d2, ax * ax + ay * ay);
                    }
                }
                const double eff = d2 < 25 ? c_repel[d2] : 0.0;
                return dp + eff;
            };
            // LAZY NOISE.  random.uniform(-0.1, 0.1) (people.py:290) can only decide between directions whose deterministic
            // scores lie within 0.2 of the best one.  Pass A evaluates the deterministic scores of every mover; a mover with a
            // single such contender has its direction without any draw (the noise of direction d is keyed by d, so not
            // drawing the others changes nothing).  Movers with two or more contenders (ties of the floor field, neighbours
            // of the best cell taken) are listed per warp and get the full keyed evaluation in pass B.  The list reuses the
            // hurt lists of phase 1 (dead since the barrier): (N + T) / worker-warps entries per warp >= its movers.
            constexpr double NOISE_MARGIN = 0.25;                  // > 0.1 - (-0.1) plus any rounding of (det + noise)
            constexpr int WW = TW / 32;
            uint16_t* const my_amb = sm.hurt + (size_t)(wt >> 5) * (size_t)(((N + T) / WW) & ~7);
            int n_amb = 0;                                          // warp-uniform
            // U items per lane and iteration, written stage by stage and branch-free so that the U independent chains interleave
            constexpr int U = (WPE >= 8 && !BIG) ? MQ_SCORE_U8 : SCORE_UNROLL;
            for (int it0 = wt; it0 < n_items; it0 += U * TW) {
                int mi[U], x[U], y[U];
                bool act[U];
                double dpv[U][2];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int it = it0 + u * TW;
                    mi[u] = it >> 2;
                    act[u] = it < n_items && mi[u] < n_mov;
                    const int i = act[u] ? (int)sm.mov[mi[u]] : 0;
                    const uint32_t p = act[u] ? sm.pos[i] : 0x00010001u;       // (1,1): neighbours stay inside the bitmap
                    x[u] = (int)(p & 0xFFFFu); y[u] = (int)(p >> 16);
                    dpv[u][0] = -INFINITY; dpv[u][1] = -INFINITY;
                    if (act[u]) {
                        const double2 v = __ldg(reinterpret_cast<const double2*>(lay.dp5 + (size_t)(x[u] * stride + y[u]) * 8) + q);
                        dpv[u][0] = v.x; dpv[u][1] = v.y;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bool a0, a1;
                    const double s0 = det_score(x[u], y[u], dpv[u][0], q * 2, a0), s1 = det_score(x[u], y[u], dpv[u][1], q * 2 + 1, a1);
                    double gmax = fmax(a0 ? s0 : -INFINITY, a1 ? s1 : -INFINITY);
                    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 1));
                    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 2));
                    const bool c0 = a0 && s0 >= gmax - NOISE_MARGIN, c1 = a1 && s1 >= gmax - NOISE_MARGIN;
                    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, c0), b1 = __ballot_sync(0xFFFFFFFFu, c1);
                    const int sh = lane & 28;
                    const int cnt = __popc((b0 >> sh) & 0xFu) + __popc((b1 >> sh) & 0xFu);
                    if (act[u] && cnt == 1 && (c0 || c1)) sm.mv[mi[u]] = (uint32_t)(q * 2 + (c1 ? 1 : 0)) << 20;
                    const bool amb = act[u] && cnt >= 2 && q == 0;
                    const uint32_t ab = __ballot_sync(0xFFFFFFFFu, amb);
                    if (amb) my_amb[n_amb + __popc(ab & ((1u << lane) - 1u))] = (uint16_t)mi[u];
                    n_amb += __popc(ab);
                }
            }
            __syncwarp();
#ifdef MQ_ENV_TRACE
            if (lane == 0) { atomicAdd((unsigned long long*)&g_env_trace[12], (unsigned long long)n_amb); }
            if (wt == 0) { atomicAdd((unsigned long long*)&g_env_trace[13], (unsigned long long)n_mov); }
#endif
            // pass B: the listed movers, all eight directions with their keyed draws, strict '>' in direction order
            // (people.py:293): larger score wins, ties go to the lower direction
            for (int e0 = 0; e0 < n_amb * 4; e0 += 32 * U) {
                int mi[U], i[U], x[U], y[U];
                bool act[U];
                double dpv[U][2];
                uint4 w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = (e0 + 32 * u + lane) >> 2;
                    act[u] = e < n_amb;
                    mi[u] = act[u] ? (int)my_amb[e] : 0;
                    i[u] = act[u] ? (int)sm.mov[mi[u]] : 0;
                    const uint32_t p = act[u] ? sm.pos[i[u]] : 0x00010001u;
                    x[u] = (int)(p & 0xFFFFu); y[u] = (int)(p >> 16);
                    dpv[u][0] = -INFINITY; dpv[u][1] = -INFINITY;
                    if (act[u]) {
                        const double2 v = __ldg(reinterpret_cast<const double2*>(lay.dp5 + (size_t)(x[u] * stride + y[u]) * 8) + q);
                        dpv[u][0] = v.x; dpv[u][1] = v.y;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) w[u] = philox4x32(env_id, tick, (uint32_t)i[u], (uint32_t)q, cfg.pk);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double best_score = -INFINITY;
                    int best_dir = 8;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        bool adm;
                        const double det = det_score(x[u], y[u], dpv[u][k], q * 2 + k, adm);
                        const double un = k ? u53(w[u].z, w[u].w) : u53(w[u].x, w[u].y);
                        const double noise = -0.1 + (0.1 - -0.1) * un;       // random.uniform(-0.1, 0.1)
                        const double score = det + noise;                      // people.py:287-291
                        if (adm && score > best_score) { best_score = score; best_dir = q * 2 + k; }
                    }
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        const double os = __shfl_xor_sync(0xFFFFFFFFu, best_score, o);
                        const int od = __shfl_xor_sync(0xFFFFFFFFu, best_dir, o);
                        if (os > best_score || (os == best_score && od < best_dir)) { best_score = os; best_dir = od; }
                    }
                    if (act[u] && q == 0 && best_dir < 8) sm.mv[mi[u]] = (uint32_t)best_dir << 20;
                }
            }
        }
        ENV_MARK(2);      // scoring
        {
            const int n_items = (n_mov * 4 + 31) & ~31;
            // proposals, one mover per lane: the movers this warp has just scored (8 per scoring iteration), so only a
            // __syncwarp separates the two loops.  move_plan[(new_x,new_y)] (people.py:228-230) + shuffle priority (:239)
            __syncwarp();
            const int w0 = (wt & ~31);                            // first worker thread id of this warp
            for (int itb = w0; itb < n_items; itb += 4 * TW) {
                const int mi = ((itb + (lane >> 3) * TW) >> 2) + (lane & 7);
                const bool in_iter = itb + (lane >> 3) * TW < n_items;
                if (!in_iter || mi >= n_mov) continue;
                const uint32_t mv = sm.mv[mi];
                if (mv == 0xFFFFFFFFu) continue;
                const int i = sm.mov[mi];
                const int best_dir = (int)(mv >> 20);
                const uint32_t p = sm.pos[i];
                const int x = (int)(p & 0xFFFFu), y = (int)(p >> 16);
                const uint4 w4 = philox4x32(env_id, tick, (uint32_t)i, STREAM_HEALTH, cfg.pk);
                const uint32_t t = (uint32_t)((x + move_dx(best_dir)) * stride + (y + move_dy(best_dir)));
                uint32_t hh = hash_cell(t, hcap);
                for (;;) {
                    const uint32_t prev = atomicCAS(&sm.tab[hh].key, HEMPTY, t);
                    if (prev == HEMPTY || prev == t) break;
                    hh = hh + 1 == hcap ? 0u : hh + 1;
                }
                atomicMin(&sm.tab[hh].ml, (uint32_t)i);                      // leave half is still 0
                atomicMin(&sm.tab[hh].best, ((unsigned long long)w4.z << 32) | (unsigned long long)i);
                { const uint32_t fb = filt_bit(t, sm.tfilt_bits); atomicOr(&sm.tfilt[fb >> 5], 1u << (fb & 31)); }
                sm.mv[mi] = hh | ((uint32_t)best_dir << 20);
            }
        }
        g.wsync();

        ENV_MARK(3);      // proposals
        // ---- phase 4a: winners leave their old cell (people.py:239-246,301) ---------------------------------
        for (int mi = wt; mi < n_mov; mi += TW) {
            const uint32_t mv = sm.mv[mi];
            if (mv == 0xFFFFFFFFu) continue;
            const int i = sm.mov[mi];
            const uint32_t slot = mv & 0xFFFFFu;
            if ((uint32_t)tab_ld<BIG>(&sm.tab[slot].best) != (uint32_t)i) continue;      // lost the shuffle: stays (people.py:248-249)
            sm.mv[mi] = mv | 0x80000000u;
            const uint32_t key = tab_ld<BIG>(&sm.tab[slot].ml) & 0xFFFFu;
            const uint32_t p = sm.pos[i];
            const int x = (int)(p & 0xFFFFu), y = (int)(p >> 16);
            const uint32_t c_old = (uint32_t)(x * stride + y);
            uint32_t hh = hash_cell(c_old, hcap);
            bool found = false;
            // most old cells are nobody's target: the filter answers that without walking a probe chain of the 2/3-full table
            const uint32_t fb = filt_bit(c_old, sm.tfilt_bits);
            if ((sm.tfilt[fb >> 5] >> (fb & 31)) & 1u) {
                for (;;) {
                    const uint32_t k = tab_ld<BIG>(&sm.tab[hh].key);
                    if (k == HEMPTY) break;
                    if (k == c_old) { found = true; break; }
                    hh = hh + 1 == hcap ? 0u : hh + 1;
                }
            }
            // old cell is somebody's target: order decides (the low half of ml is final since the barrier)
            if (found) atomicMax(&sm.tab[hh].ml, ((key + 1u) << 16) | (tab_ld<BIG>(&sm.tab[hh].ml) & 0xFFFFu));
            else bm_clear(sm.bm, wpr, x, y);                    // rmap[old] = 0
        }
        g.wsync();

        // ---- phase 4b: winners enter their target (people.py:302-314) --------------------------------------
        for (int mi = wt; mi < n_mov; mi += TW) {
            const uint32_t mv = sm.mv[mi];
            if (mv == 0xFFFFFFFFu || !(mv & 0x80000000u)) continue;
            const int i = sm.mov[mi];
            const uint32_t ml = tab_ld<BIG>(&sm.tab[mv & 0xFFFFFu].ml);
            const int dir = (int)((mv >> 20) & 7u);
            const uint32_t p = sm.pos[i];
            const int nx = (int)(p & 0xFFFFu) + move_dx(dir), ny = (int)(p >> 16) + move_dy(dir);
            const bool evac = (__ldg(lay.cellinfo + nx * stride + ny) & 8u) != 0;      // Map.checkSavefy (map.py:93-113)
            if ((ml & 0xFFFFu) + 1u > (ml >> 16) && !evac) bm_set(sm.bm, wpr, nx, ny);
            else bm_clear(sm.bm, wpr, nx, ny);
            const uint32_t np = (uint32_t)nx | ((uint32_t)ny << 16);
            sm.pos[i] = np;
            st.pos[base + i] = np;
            if (evac) { const uint8_t f = sm.fl[i] | 1u; sm.fl[i] = f; st.flags[base + i] = f; }
        }
        g.wsync();      // the proposal table is dead from here on: sm.dist may overwrite it

        ENV_MARK(4);      // moves
        // ---- reward inputs (evacuation_env.py:174-233): counts, guidance, ordered list of distances --------
        const int rpx = sc[MQ_S_ROBOT_POS_X], rpy = sc[MQ_S_ROBOT_POS_Y];
        {
            const int P = (N + TW - 1) / TW;
            const int i0 = min(wt * P, N), i1 = min(i0 + P, N);
            int live = 0, evac = 0, dead = 0, guid = 0;
            for (int i = i0; i < i1; ++i) {
                const uint32_t f = sm.fl[i];
                evac += f & 1u; dead += (f >> 1) & 1u;
                if (!(f & 3u)) ++live;
            }
            // exclusive scan of `live` over the workers
            int incl = live;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
            int off = incl - live;
            if (WPE > 1) {
                if (lane == 31) s_wtot[warp] = incl;
                g.wsync();
                for (int w = 0; w < warp; ++w) off += s_wtot[w];
            }
            for (int i = i0; i < i1; ++i) {
                const uint32_t f = sm.fl[i];
                if (f & 3u) continue;
                const uint32_t p = sm.pos[i];
                const int x = (int)(p & 0xFFFFu), y = (int)(p >> 16);
                // 2*(x+.5-rx) is an odd integer: compare squared distances exactly (sqrt is monotone, thresholds exact)
                const long long X = 2LL * x + 1 - 2LL * rpx, Y = 2LL * y + 1 - 2LL * rpy;
                const long long q4 = X * X + Y * Y;                             // 4 * dist^2
                if (q4 <= 100) {                                                // dist_to_robot <= 5 (:202)
                    const long long EX = 2LL * x + 1 - 2LL * lay.obs_exit[0], EY = 2LL * y + 1 - 2LL * lay.obs_exit[1];
                    const long long e4 = EX * EX + EY * EY;
                    guid += e4 > 1600 ? 4 : (e4 > 400 ? 3 : 2);                 // 2.0 / 1.5 / 1.0 (:207-212)
                    if ((HSM ? sm.health[i] : __ldcg(st.health + base + i)) < 80.0) guid += 2;      // +1.0 (:215-216)
                }
                const double dx = ((double)x + 0.5) - (double)rpx, dy = ((double)y + 0.5) - (double)rpy;
                sm.dist[off++] = sqrt(dx * dx + dy * dy);                       // np.linalg.norm (:228)
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                evac += __shfl_xor_sync(0xFFFFFFFFu, evac, o);
                dead += __shfl_xor_sync(0xFFFFFFFFu, dead, o);
                guid += __shfl_xor_sync(0xFFFFFFFFu, guid, o);
                live += __shfl_xor_sync(0xFFFFFFFFu, live, o);
            }
            if (lane == 0) {
                if (WPE == 1) { s_cnt[0] = evac; s_cnt[1] = dead; s_cnt[2] = guid; s_cnt[3] = live; }
                else { atomicAdd(&s_cnt[0], evac); atomicAdd(&s_cnt[1], dead); atomicAdd(&s_cnt[2], guid); atomicAdd(&s_cnt[3], live); }
            }
        }
        g.wsync();

        ENV_MARK(5);      // reward inputs
        // ---- np.mean tree (numpy pairwise summation, evacuation_env.py:228) --------------------------------------------
        if (WPE >= 8) {
            // Wide groups: the serial enumerate / combine of the leaves cost 160 us per step at 20,000 people.  Warp 0 builds the
            // tree level by level (node arrays in level order, children found with a ballot scan) while the other workers
            // write the outputs; every 8-lane group of the group then sums leaves; warp 0 adds the levels bottom-up.
            int* lvl = s_lvl_all[WPE >= 8 ? g.gid : 0];
            const int fs_new = min(sc[MQ_S_FIRE_STEP] + 1, lay.n_fire_steps - 1);
            if (warp == 0) {
                const int n = s_cnt[3];
                if (lane == 0) { sm.leaf_off[0] = 0; sm.leaf_len[0] = n; }
                __syncwarp();
                int lo = 0, hi = 1, nlev = 0;
                for (;;) {
                    if (lane == 0) lvl[nlev] = lo;
                    ++nlev;
                    int next = hi;
                    for (int b0 = lo; b0 < hi; b0 += 32) {
                        const int i = b0 + lane;
                        const int len = i < hi ? sm.leaf_len[i] : 0;
                        const bool split = len > 128;
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, split);
                        const int c = next + 2 * __popc(bal & ((1u << lane) - 1u));
                        if (i < hi) sm.nchild[i] = split ? c : -1;
                        if (split) {
                            const int off = sm.leaf_off[i];
                            int n2 = len / 2; n2 -= n2 % 8;
                            sm.leaf_off[c] = off; sm.leaf_len[c] = n2;
                            sm.leaf_off[c + 1] = off + n2; sm.leaf_len[c + 1] = len - n2;
                        }
                        next += 2 * __popc(bal);
                    }
                    __syncwarp();
                    if (next == hi) break;          // nothing split: [lo, hi) is the last level
                    lo = hi; hi = next;
                }
                if (lane == 0) { lvl[nlev] = hi; lvl[19] = nlev; }
            } else {
                if (n_mov > 0) store_bitmap(wt - 32, TW - 32, lay, sm, st.rmap, env);      // nobody moved: the global copy is current
                gather_obs(wt - 32, TW - 32, lay, cfg, sm, rob, rpx, rpy, fs_new, obs, obs64, env);
            }
            g.wsync();
            {
                const int nlev = lvl[19], n_nodes = lvl[nlev];
                const int gl = lane & 7;
                const uint32_t gmask = 0xFFu << ((lane >> 3) * 8);
                for (int nd = wt >> 3; nd < n_nodes; nd += TW >> 3) {         // one 8-lane group per node
                    if (sm.nchild[nd] < 0) {
                        const double sv = leaf_sum8(sm.dist + sm.leaf_off[nd], sm.leaf_len[nd], gl, gmask);
                        if (gl == 0) sm.leaf_sum[nd] = sv;
                    }
                }
            }
            g.wsync();
            if (warp == 0) {
                const int nlev = lvl[19];
                for (int L = nlev - 2; L >= 0; --L) {                          // the last level holds leaves only
                    for (int i = lvl[L] + lane; i < lvl[L + 1]; i += 32) {
                        const int c = sm.nchild[i];
                        if (c >= 0) sm.leaf_sum[i] = sm.leaf_sum[c] + sm.leaf_sum[c + 1];      // pairwise(left) + pairwise(right)
                    }
                    __syncwarp();
                }
                if (lane == 0) s_sum[0] = sm.leaf_sum[0];
                if (!G::CHAIN) {
                    const double th = HSM ? health_chain_runs<BIG ? 8 : 1>(sm.health, N, lane)
                                          : health_chain_runs<4, true>(st.health + base, N, lane, sm.fl);
                    if (lane == 0) s_sum[1] = th;
                }
            }
        } else {
        if (warp == 0) {
            const int n = s_cnt[3];
            int nl = 0;
            if (lane == 0) nl = enumerate_leaves(n, sm.leaf_off, sm.leaf_len);
            nl = __shfl_sync(0xFFFFFFFFu, nl, 0);
            __syncwarp();
            const int gl = lane & 7, grp = lane >> 3;
            const uint32_t gmask = 0xFFu << (grp * 8);
            for (int l0 = 0; l0 < nl; l0 += 4) {
                const int l = l0 + grp;
                if (l < nl) {
                    const double s = leaf_sum8(sm.dist + sm.leaf_off[l], sm.leaf_len[l], gl, gmask);
                    if (gl == 0) sm.leaf_sum[l] = s;
                }
            }
            __syncwarp();
            if (lane == 0) s_sum[0] = combine_leaves(n, sm.leaf_sum);
            if (!G::CHAIN) {
                if (WPE == 1) { if (lane == 0) s_sum[1] = health_chain_literal(sm.health, N); }
                else { const double th = health_chain_runs<BIG ? 8 : 1>(sm.health, N, lane); if (lane == 0) s_sum[1] = th; }
            }
        }
        ENV_MARK(6);      // tree + chain
        // fire models step (evacuation_env.py:138-142; fire_model.py:63-67) — the observation uses the new step
        if (n_mov > 0) store_bitmap(wt, TW, lay, sm, st.rmap, env);      // nobody moved: the global copy is current
        gather_obs(wt, TW, lay, cfg, sm, rob, rpx, rpy, min(sc[MQ_S_FIRE_STEP] + 1, lay.n_fire_steps - 1), obs, obs64, env);
        }
    }
    g.sync();

    ENV_MARK(7);      // bitmap + obs
    // ---- _calculate_reward (evacuation_env.py:174-288), done (:150-157), scalars -----------------------
    if (tid == 0) {
        const int cur_evac = s_cnt[0], cur_dead = s_cnt[1], n_rem = s_cnt[3];
        const int cur_step = sc[MQ_S_CUR_STEP];
        const int remaining = N - cur_evac - cur_dead;
        const double total_health = s_sum[1];
        double reward = 0.0;
        reward += (double)(cur_evac - sc[MQ_S_PREV_EVAC]) * cfg.evac_reward;
        reward += (double)s_cnt[2] * 0.5;
        if (remaining > 0 && n_rem > 0) {
            const double avg = s_sum[0] / (double)n_rem;
            double dr = 2.0 - fabs(avg - 8.0) * 0.2;
            if (!(dr > 0.0)) dr = 0.0;
            reward += dr;
        }
        if (remaining > 0) {
            const double urgency = (double)remaining / (double)N;
            const double time_penalty = -0.05 - (urgency * 0.1);
            reward += time_penalty;
        } else {
            reward -= 0.02;
        }
        if (N - cur_dead > 0) {
            const double avg_health = total_health / (double)(N - cur_dead);
            reward += (avg_health - 90.0) * 0.05;
        }
        if (cur_evac == N) {
            const int rem_steps = max(0, 300 - cur_step);
            const double time_bonus = (double)rem_steps * 0.2;
            const double final_avg = total_health / (double)N;      // nobody is dead here: same chain
            const double health_bonus = (final_avg - 80.0) * 1.0;
            reward += (100.0 + time_bonus) + health_bonus;
        }
        reward -= (double)(cur_dead - sc[MQ_S_PREV_DEAD]) * cfg.death_penalty;
        reward -= (double)cur_dead * cfg.death_acc_penalty;
        reward += (double)(N - cur_dead) * cfg.alive_bonus;
        if (cur_step > 0) {
            const double eff = (double)cur_evac / (double)cur_step;
            if (eff > 0.1) reward += eff * 5.0;
        }
        const int done = (cur_evac + cur_dead == N) || (cur_step + 1 >= cfg.max_steps);
        reward_out[env] = reward;
        done_out[env] = (uint8_t)done;
        sc[MQ_S_FIRE_STEP] = min(sc[MQ_S_FIRE_STEP] + 1, lay.n_fire_steps - 1);
        sc[MQ_S_CUR_STEP] = cur_step + 1;
        sc[MQ_S_PREV_EVAC] = cur_evac; sc[MQ_S_PREV_DEAD] = cur_dead;
        sc[MQ_S_EVAC] = cur_evac; sc[MQ_S_DEAD] = cur_dead;
        sc[MQ_S_TICK] = (int)(tick + 1u);
        s_cnt[5] = done && cfg.auto_reset;
    }
    g.sync();
    ENV_MARK(8);      // reward
    if (s_cnt[5]) {
        reset_env<WPE, CW>(g, lay, cfg, st, sm, sc, rob, nullptr, obs, obs64, env);
    } else {
        if (tid < MQ_ENV_SCALARS) st.scalars[(size_t)env * MQ_ENV_SCALARS + tid] = sc[tid];
        if (tid < MAXR * 2) st.robots[(size_t)env * MAXR * 2 + tid] = rob[tid >> 1][tid & 1];
    }
}

__global__ void unpack_rmap_kernel(DevLayout lay, const uint32_t* __restrict__ rmap, uint8_t* __restrict__ out,
                                   long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    long long env = idx / lay.G;
    int c = (int)(idx - env * lay.G);
    int x = c / lay.stride, y = c - x * lay.stride;
    out[idx] = (uint8_t)((rmap[env * lay.rmap_words + x * lay.wpr + (y >> 5)] >> (y & 31)) & 1u);
}

}  // namespace mq

// =================================================================================================
// C-ABI
// =================================================================================================
struct mq_env {
    mq::DevLayout lay;
    mq::DevCfg cfg;
    mq::DevState st;
    void* d_dp5 = nullptr; void* d_cellinfo = nullptr; void* d_ctr = nullptr; void* d_int = nullptr;
    void* d_meta = nullptr; void* d_env_layout = nullptr;
    int n_layouts = 1;
    int device = 0;
    int wpe = 8;            // warps per env: 1 (4 envs per 128-thread CTA), 8 (one 256-thread CTA per env), 16 / 32 (BIG)
    bool big = false;       // per-person arrays + proposal table in global scratch (envs too large for shared memory)
    int variant = 0;        // index into the kernel variant table
    void* d_scratch = nullptr;
    int blocks = 0, threads = 0;
    size_t smem = 0;        // dynamic shared memory per CTA
    int64_t launches = 0;
};

static int round_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
constexpr int SMALL_CW = 4;      // warps per CTA of the small-env variants (1, 2 or 4 warps per env)
constexpr int SMALL_WPE = 1;     // default warps per env for small envs
constexpr int BIG_WPE = 32;      // warps per env (= per CTA) of the BIG variant

extern "C" int mq_env_state_sizes(const mq_env_cfg* cfg, const mq_layout* layout, int64_t* n_pad, int64_t* rmap_words) {
    MQ_REQUIRE(cfg && layout, "mq_env_state_sizes: null argument");
    if (n_pad) *n_pad = (cfg->n_people + 15) / 16 * 16;
    if (rmap_words) {
        int64_t wpr = (layout->W + 2 + 31) / 32;
        *rmap_words = ((int64_t)(layout->L + 2) * wpr + 3) / 4 * 4;
    }
    return MQ_OK;
}

// kernel variants: (warps per env, warps per CTA, BIG)
typedef void (*StepFn)(mq::DevLayout, mq::DevCfg, mq::DevState, const int*, float*, double*, double*, uint8_t*);
typedef void (*ResetFn)(mq::DevLayout, mq::DevCfg, mq::DevState, const uint8_t*, const int16_t*, float*, double*);
struct Variant { int wpe, cw; bool big, multi; StepFn step; ResetFn reset; };
#define MQ_VARIANT(WPE, CW, BIG, MULTI) {WPE, CW, BIG, MULTI, mq::env_step_kernel<WPE, CW, BIG, MULTI>, mq::env_reset_kernel<WPE, CW, BIG, MULTI>}
static const Variant k_variants[] = {
    MQ_VARIANT(1, SMALL_CW, false, false), MQ_VARIANT(1, 14, false, false), MQ_VARIANT(1, 28, false, false), MQ_VARIANT(2, SMALL_CW, false, false),
    MQ_VARIANT(4, SMALL_CW, false, false), MQ_VARIANT(8, 8, false, false), MQ_VARIANT(8, 8, true, false), MQ_VARIANT(16, 16, true, false),
    MQ_VARIANT(32, 32, true, false),
    // several layouts per batch: the shapes mq_env_create picks by default
    MQ_VARIANT(1, SMALL_CW, false, true), MQ_VARIANT(1, 14, false, true), MQ_VARIANT(1, 28, false, true), MQ_VARIANT(8, 8, false, true),
    MQ_VARIANT(32, 32, true, true),
};
static int find_variant(int wpe, bool big, bool multi, int cw = 0) {
    for (int i = 0; i < (int)(sizeof(k_variants) / sizeof(k_variants[0])); ++i)
        if (k_variants[i].wpe == wpe && k_variants[i].big == big && k_variants[i].multi == multi && (cw == 0 || k_variants[i].cw == cw)) return i;
    return -1;
}

extern "C" int mq_env_create_layouts(mq_env** out, const mq_env_cfg* cfg, const mq_layout* layouts, int32_t n_layouts,
                                     const int32_t* env_layout, int32_t tables_on_device, const mq_env_state* state) {
    MQ_REQUIRE(out && cfg && layouts && state && n_layouts >= 1, "mq_env_create: null argument");
    MQ_REQUIRE(n_layouts == 1 || env_layout, "mq_env_create_layouts: env_layout is required when there are several layouts");
    const mq_layout* layout = &layouts[0];
    MQ_REQUIRE(cfg->n_envs > 0 && cfg->n_people > 0, "mq_env_create: n_envs and n_people must be positive");
    MQ_REQUIRE(cfg->n_robots >= 1 && cfg->n_robots <= MQ_MAX_ROBOTS, "mq_env_create: n_robots must be in 1..%d", MQ_MAX_ROBOTS);
    MQ_REQUIRE(layout->L >= 3 && layout->W >= 3 && layout->L <= 32000 && layout->W <= 32000, "mq_env_create: bad grid size");
    for (int k = 0; k < n_layouts; ++k) {
        MQ_REQUIRE(layouts[k].dp5 && layouts[k].cellinfo && layouts[k].danger_ctr && layouts[k].danger_int, "mq_env_create: tables of layout %d missing", k);
        MQ_REQUIRE(layouts[k].L == layout->L && layouts[k].W == layout->W && layouts[k].n_fire_steps == layout->n_fire_steps,
                   "mq_env_create_layouts: layout %d differs in L x W or fire steps from layout 0", k);
    }
    if (n_layouts > 1)
        for (int e = 0; e < cfg->n_envs; ++e)
            MQ_REQUIRE(env_layout[e] >= 0 && env_layout[e] < n_layouts, "mq_env_create_layouts: env_layout[%d] = %d outside 0..%d", e, env_layout[e], n_layouts - 1);
    MQ_REQUIRE(state->pos && state->health && state->acc && state->flags && state->rmap && state->robots && state->scalars,
               "mq_env_create: state buffers missing");
    MQ_REQUIRE(cfg->n_people <= 65000, "mq_env_create: at most 65000 people per env (16-bit list indices)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return mq::fail(MQ_ERR_CUDA, "mq_env_create: no CUDA device (this build has no CPU fallback)");
    MQ_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "mq_env_create: device %d outside 0..%d", cfg->device, ndev - 1);
    MQ_ON_DEVICE(cfg->device);

    mq_env* e = new (std::nothrow) mq_env();
    if (!e) return mq::fail(MQ_ERR_ALLOC, "mq_env_create: out of host memory");
    e->device = cfg->device;
    mq::DevLayout& l = e->lay;
    l.L = layout->L; l.W = layout->W; l.stride = layout->W + 2; l.G = (layout->L + 2) * (layout->W + 2);
    l.wpr = (layout->W + 2 + 31) / 32;
    l.rmap_words = ((layout->L + 2) * l.wpr + 3) / 4 * 4;
    l.n_fire_steps = layout->n_fire_steps;
    memcpy(l.ctr_box, layout->ctr_box, sizeof(l.ctr_box));
    memcpy(l.int_box, layout->int_box, sizeof(l.int_box));
    memcpy(l.robot_range, layout->robot_range, sizeof(l.robot_range));
    memcpy(l.robot_start, layout->robot_start, sizeof(l.robot_start));
    memcpy(l.reset_center, layout->reset_obs_center, sizeof(l.reset_center));
    memcpy(l.obs_exit, layout->obs_exit, sizeof(l.obs_exit));
    l.env_layout = nullptr; l.meta = nullptr;

    // Tables of all layouts, contiguous per kind.  Danger tables that several layouts share (same source pointer and size: the
    // usual case of random walls around one fire) are stored once.
    const size_t n_dp5 = (size_t)l.G * 8 * sizeof(double);
    std::vector<int> meta((size_t)n_layouts * mq::LMETA, 0);
    std::vector<size_t> ctr_off(n_layouts), int_off(n_layouts), ctr_n(n_layouts), int_n(n_layouts);
    size_t ctr_total = 0, int_total = 0;
    for (int k = 0; k < n_layouts; ++k) {
        const mq_layout& lk = layouts[k];
        ctr_n[k] = (size_t)lk.n_fire_steps * lk.ctr_box[2] * lk.ctr_box[3];
        int_n[k] = (size_t)lk.n_fire_steps * lk.int_box[2] * lk.int_box[3];
        int dup_c = -1, dup_i = -1;
        for (int q = 0; q < k && (dup_c < 0 || dup_i < 0); ++q) {
            if (dup_c < 0 && layouts[q].danger_ctr == lk.danger_ctr && ctr_n[q] == ctr_n[k]) dup_c = q;
            if (dup_i < 0 && layouts[q].danger_int == lk.danger_int && int_n[q] == int_n[k]) dup_i = q;
        }
        if (dup_c >= 0) ctr_off[k] = ctr_off[dup_c]; else { ctr_off[k] = ctr_total; ctr_total += ctr_n[k]; }
        if (dup_i >= 0) int_off[k] = int_off[dup_i]; else { int_off[k] = int_total; int_total += int_n[k]; }
        int* m = &meta[(size_t)k * mq::LMETA];
        memcpy(m + mq::LM_CTR_BOX, lk.ctr_box, 4 * sizeof(int)); memcpy(m + mq::LM_INT_BOX, lk.int_box, 4 * sizeof(int));
        memcpy(m + mq::LM_ROBOT_RANGE, lk.robot_range, 2 * sizeof(int)); memcpy(m + mq::LM_ROBOT_START, lk.robot_start, 8 * sizeof(int));
        memcpy(m + mq::LM_RESET_CENTER, lk.reset_obs_center, 2 * sizeof(int)); memcpy(m + mq::LM_OBS_EXIT, lk.obs_exit, 2 * sizeof(int));
        const long long co = (long long)ctr_off[k], io = (long long)int_off[k];
        memcpy(m + mq::LM_CTR_OFF, &co, sizeof(co)); memcpy(m + mq::LM_INT_OFF, &io, sizeof(io));
    }
    cudaError_t ce = cudaSuccess;
    auto step = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
    step(cudaMalloc(&e->d_dp5, n_dp5 * n_layouts));
    step(cudaMalloc(&e->d_cellinfo, (size_t)l.G * n_layouts));
    step(cudaMalloc(&e->d_ctr, ctr_total * sizeof(double)));
    step(cudaMalloc(&e->d_int, int_total * sizeof(double)));
    const cudaMemcpyKind kind = tables_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    for (int k = 0; k < n_layouts && ce == cudaSuccess; ++k) {
        const mq_layout& lk = layouts[k];
        step(cudaMemcpy((char*)e->d_dp5 + n_dp5 * k, lk.dp5, n_dp5, kind));
        step(cudaMemcpy((char*)e->d_cellinfo + (size_t)l.G * k, lk.cellinfo, l.G, kind));
        bool first_c = true, first_i = true;
        for (int q = 0; q < k; ++q) { if (ctr_off[q] == ctr_off[k]) first_c = false; if (int_off[q] == int_off[k]) first_i = false; }
        if (first_c) step(cudaMemcpy((double*)e->d_ctr + ctr_off[k], lk.danger_ctr, ctr_n[k] * sizeof(double), kind));
        if (first_i) step(cudaMemcpy((double*)e->d_int + int_off[k], lk.danger_int, int_n[k] * sizeof(double), kind));
    }
    if (n_layouts > 1) {
        step(cudaMalloc(&e->d_meta, meta.size() * sizeof(int)));
        step(cudaMalloc(&e->d_env_layout, (size_t)cfg->n_envs * sizeof(int)));
        if (ce == cudaSuccess) step(cudaMemcpy(e->d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (ce == cudaSuccess) step(cudaMemcpy(e->d_env_layout, env_layout, (size_t)cfg->n_envs * sizeof(int), cudaMemcpyHostToDevice));
        l.meta = (const int*)e->d_meta; l.env_layout = (const int*)e->d_env_layout;
    }
    if (ce != cudaSuccess) {
        mq_env_destroy(e);
        return mq::fail(MQ_ERR_CUDA, "mq_env_create: uploading layout tables failed: %s", cudaGetErrorString(ce));
    }
    l.dp5 = (const double*)e->d_dp5; l.cellinfo = (const uint8_t*)e->d_cellinfo;
    l.danger_ctr = (const double*)e->d_ctr; l.danger_int = (const double*)e->d_int;
    e->n_layouts = n_layouts;

    double repel[25];
    for (int d2 = 0; d2 < 25; ++d2) repel[d2] = -20.0 / (std::sqrt((double)d2) + 0.1);    // people.py:94,284
    ce = cudaMemcpyToSymbol(mq::c_repel, repel, sizeof(repel));
    if (ce != cudaSuccess) { mq_env_destroy(e); return mq::fail(MQ_ERR_CUDA, "mq_env_create: c_repel upload: %s", cudaGetErrorString(ce)); }

    mq::DevCfg& c = e->cfg;
    c.n_envs = cfg->n_envs; c.N = cfg->n_people; c.n_pad = (cfg->n_people + 15) / 16 * 16; c.R = cfg->n_robots;
    c.seed = cfg->seed; c.pk = mq::philox_keys(cfg->seed); c.env_id_base = cfg->env_id_base; c.max_steps = cfg->max_steps;
    c.reset_robots = cfg->reset_robots; c.reset_fire = cfg->reset_fire; c.auto_reset = cfg->auto_reset;
    int want = c.N + c.N / 3 + 8;                      // load factor <= 0.75 even if everybody proposes a distinct cell
    c.hash_cap = round_pow2(want < 64 ? 64 : want);    // < 2^20 (20-bit slot ids in Smem::mv)
    if (c.hash_cap >= 128 && c.hash_cap / 4 * 3 >= want) c.hash_cap = c.hash_cap / 4 * 3;      // 1.5 x 2^k is enough: 1536 slots (24 KB) at 1000 people
    c.health_smem = 1;
    if (const char* v = getenv("MQ_ENV_HSM")) c.health_smem = atoi(v) != 0;                    // A/B timing knobs
    if (const char* v = getenv("MQ_ENV_HASH_POW2")) { if (atoi(v)) c.hash_cap = round_pow2(want < 64 ? 64 : want); }
    c.n_leaf_max = c.N / 64 + 4;
    c.evac_reward = cfg->evac_reward; c.death_penalty = cfg->death_penalty;
    c.death_acc_penalty = cfg->death_acc_penalty; c.alive_bonus = cfg->alive_bonus;
    c.obs_wire = nullptr;
    e->st = {state->pos, state->health, state->acc, state->flags, state->rmap, state->robots, state->scalars};

    mq::Smem tmp;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    c.scratch = nullptr; c.scratch_per_env = 0;
    c.smem_per_env = (int)mq::carve(tmp, nullptr, nullptr, c.N, c.hash_cap, l.rmap_words, c.n_leaf_max, 32);
    // one warp per env for small envs (no CTA barriers, 4 envs per CTA), one CTA per env otherwise; envs whose
    // person arrays do not fit shared memory keep only the occupancy bitmap there (BIG)
    e->wpe = (c.N <= 256 && (size_t)c.smem_per_env * SMALL_CW + 2048 <= (size_t)max_smem) ? SMALL_WPE : 8;
    if (e->wpe != 8) {
        if (const char* v = getenv("MQ_SMALL_WPE")) { int w = atoi(v); if (w == 1 || w == 2 || w == 4) e->wpe = w; }
        c.smem_per_env = (int)mq::carve(tmp, nullptr, nullptr, c.N, c.hash_cap, l.rmap_words, c.n_leaf_max, 32 * e->wpe);
    }
    if (e->wpe == 8) c.smem_per_env = (int)mq::carve(tmp, nullptr, nullptr, c.N, c.hash_cap, l.rmap_words, c.n_leaf_max, 256, nullptr, c.health_smem != 0);
    e->big = e->wpe == 8 && (size_t)c.smem_per_env + 2048 > (size_t)max_smem;
    if (e->big) {
        e->wpe = BIG_WPE;                              // one CTA per SM (shared-memory bound): make it a wide one
        if (const char* v = getenv("MQ_BIG_WPE")) { int w = atoi(v); if (w == 8 || w == 16 || w == 32) e->wpe = w; }
        size_t gbytes = 0;
        c.smem_per_env = (int)mq::carve(tmp, nullptr, (unsigned char*)16, c.N, c.hash_cap, l.rmap_words, c.n_leaf_max, 32 * e->wpe, &gbytes);
        c.scratch_per_env = (long long)gbytes;
        if ((ce = cudaMalloc(&e->d_scratch, gbytes * (size_t)c.n_envs)) != cudaSuccess) {
            mq_env_destroy(e);
            return mq::fail(MQ_ERR_ALLOC, "mq_env_create: %zu B of scratch for %d large envs: %s", gbytes * (size_t)c.n_envs, c.n_envs,
                            cudaGetErrorString(ce));
        }
        c.scratch = (unsigned char*)e->d_scratch;
    }
    // warp-per-env: as many envs per CTA as fit one SM (28 = one 896-thread CTA per SM, 14 = two), so that the cooperative
    // scoring phase averages over many envs; 4 envs per CTA (no cooperation) when the envs' shared memory is too large
    int small_cw = SMALL_CW;
    if (e->wpe == 1) {
        const size_t fixed = 8192;                        // static shared memory of the widest variant + driver reserve
        if ((size_t)c.smem_per_env * 28 + fixed <= (size_t)max_smem) small_cw = 28;
        else if (((size_t)c.smem_per_env * 14 + fixed) * 2 <= (size_t)max_smem + 4096) small_cw = 14;
        if (const char* v = getenv("MQ_SMALL_CW")) { int w = atoi(v); if (w == 4 || w == 14 || w == 28) small_cw = w; }
    }
    e->variant = find_variant(e->wpe, e->big, n_layouts > 1, e->wpe == 1 ? small_cw : 0);
    if (e->variant < 0) {
        mq_env_destroy(e);
        return mq::fail(MQ_ERR_UNSUPPORTED, "mq_env_create_layouts: no multi-layout kernel for %d warps per env (MQ_*_WPE / MQ_SMALL_CW overrides are single-layout only)", e->wpe);
    }
    const int groups = e->wpe < 8 ? small_cw / e->wpe : 1;
    e->threads = 32 * k_variants[e->variant].cw;
    e->smem = (size_t)c.smem_per_env * groups;
    if (const char* v = getenv("MQ_ENV_SMEM_PAD")) { if (e->wpe == 8 && !e->big) e->smem += (size_t)atoi(v); }     // A/B: fewer resident CTAs
    if ((int)e->smem + 2048 > max_smem) {
        size_t need = e->smem;
        mq_env_destroy(e);
        return mq::fail(MQ_ERR_UNSUPPORTED,
                        "mq_env_create: the occupancy bitmap of a %dx%d grid needs %zu B of shared memory per CTA (limit %d)",
                        layout->L, layout->W, need, max_smem);
    }
    e->blocks = (c.n_envs + groups - 1) / groups;
    // The limit is an attribute of the FUNCTION (per device), shared by every env handle that uses this variant: raise it to
    // the device's opt-in maximum once instead of to this env's size (a later, smaller env must not lower it under a live one).
    // (the dynamic part may be at most the opt-in maximum minus the kernel's static shared memory)
    for (const void* fn : {(const void*)k_variants[e->variant].step, (const void*)k_variants[e->variant].reset}) {
        cudaFuncAttributes fa{};
        ce = cudaFuncGetAttributes(&fa, fn);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - (int)fa.sharedSizeBytes);
        if (ce != cudaSuccess) break;
    }
    if (ce != cudaSuccess) {
        mq_env_destroy(e);
        return mq::fail(MQ_ERR_CUDA, "mq_env_create: cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    }
    *out = e;
    return MQ_OK;
}

extern "C" int mq_env_create(mq_env** out, const mq_env_cfg* cfg, const mq_layout* layout, const mq_env_state* state) {
    return mq_env_create_layouts(out, cfg, layout, 1, nullptr, 0, state);
}

extern "C" int mq_env_destroy(mq_env* e) {
    if (!e) return MQ_OK;
    MQ_ON_DEVICE(e->device);
    cudaFree(e->d_dp5); cudaFree(e->d_cellinfo); cudaFree(e->d_ctr); cudaFree(e->d_int); cudaFree(e->d_scratch);
    cudaFree(e->d_meta); cudaFree(e->d_env_layout);
    delete e;
    return MQ_OK;
}

extern "C" int mq_env_set_reward_coefs(mq_env* e, double evac_reward, double death_penalty, double death_acc_penalty,
                                       double alive_bonus) {
    MQ_REQUIRE(e, "mq_env_set_reward_coefs: null handle");
    e->cfg.evac_reward = evac_reward; e->cfg.death_penalty = death_penalty;
    e->cfg.death_acc_penalty = death_acc_penalty; e->cfg.alive_bonus = alive_bonus;
    return MQ_OK;
}

extern "C" int mq_env_set_obs_wire(mq_env* e, uint32_t* wire_out) {
    MQ_REQUIRE(e, "mq_env_set_obs_wire: null handle");
    e->cfg.obs_wire = wire_out;
    return MQ_OK;
}

extern "C" int mq_env_reset(mq_env* e, const uint8_t* env_mask, const int16_t* inject_spawn, float* obs_out,
                            double* obs64_out, void* stream) {
    MQ_REQUIRE(e, "mq_env_reset: null handle");
    MQ_ON_DEVICE(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    k_variants[e->variant].reset<<<e->blocks, e->threads, e->smem, s>>>(e->lay, e->cfg, e->st, env_mask, inject_spawn, obs_out, obs64_out);
    MQ_CUDA(cudaGetLastError());
    e->launches += 1;
    return MQ_OK;
}

extern "C" int mq_env_step(mq_env* e, const int32_t* actions, float* obs_out, double* obs64_out, double* reward_out,
                           uint8_t* done_out, void* stream) {
    MQ_REQUIRE(e && actions && reward_out && done_out, "mq_env_step: null argument");
    MQ_ON_DEVICE(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    k_variants[e->variant].step<<<e->blocks, e->threads, e->smem, s>>>(e->lay, e->cfg, e->st, actions, obs_out, obs64_out, reward_out, done_out);
    MQ_CUDA(cudaGetLastError());
    e->launches += 1;
    return MQ_OK;
}

#ifdef MQ_ENV_TRACE
extern "C" int mq_debug_env_trace(long long* host_out, int reset) {
    if (reset) { static long long z[16]; return (int)cudaMemcpyToSymbol(mq::g_env_trace, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(host_out, mq::g_env_trace, sizeof(long long) * 16);
}
#endif
extern "C" int mq_env_unpack_rmap(mq_env* e, uint8_t* rmap_out, void* stream) {
    MQ_REQUIRE(e && rmap_out, "mq_env_unpack_rmap: null argument");
    MQ_ON_DEVICE(e->device);
    long long total = (long long)e->cfg.n_envs * e->lay.G;
    int blocks = (int)((total + 255) / 256);
    mq::unpack_rmap_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(e->lay, e->st.rmap, rmap_out, total);
    MQ_CUDA(cudaGetLastError());
    e->launches += 1;
    return MQ_OK;
}

extern "C" int64_t mq_env_launch_count(const mq_env* e) { return e ? e->launches : 0; }
