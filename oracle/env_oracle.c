/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, sequential restatement of the reference's env hot path
 * (LX-530/DQN-MARL, Louvre_Evacuation/envs).  It is the CPU checker for the
 * CUDA kernels: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  The product
 * path (dqn_marl_b200/) never links or imports anything from oracle/.
 *
 * It deliberately follows the reference statement by statement — the dict
 * insertion order of move_plan, the flag (not count) semantics of rmap, the
 * left-to-right / pairwise fp64 summation trees of the reward — instead of the
 * parallel formulation the kernels use, so that agreement between the two is
 * evidence and not a tautology.
 *
 * Pinned: tests/test_oracle_golden.py compares it bit for bit with trajectories
 * recorded from the UNMODIFIED Python reference (oracle/make_golden.py ->
 * tests/golden/ npz files), both driven by the keyed draws of oracle/keyed_draws.py.
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXR 4
#define OBS_WIN 11
#define OBS_CH 6
#define OBS_SIZE (OBS_WIN * OBS_WIN * OBS_CH)

/* ---- keyed draws (oracle/keyed_draws.py) -------------------------------- */
static void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
enum { STREAM_HEALTH = 4, STREAM_SPAWN = 16 };

typedef struct orc_env {
    /* layout (borrowed pointers, owned by the caller) */
    int L, W, N, R;
    const double* space;        /* [G]  Map.space after Init_Potential (map.py:148) */
    const uint8_t* barrier;     /* [G]  membership in Map.barrier_list (map.py:43-57,72-73) */
    int n_exits; const int32_t* exits;  /* Map.Exit */
    int obs_exit[2];            /* EvacuationEnv.exit_location */
    int n_fire_steps;
    int ctr_box[4]; const double* danger_ctr;
    int int_box[4]; const double* danger_int;
    int robot_range[2];
    int robot_start[MAXR][2];
    int reset_obs_center[2];
    /* config */
    uint64_t seed; uint32_t env_id;
    int max_steps, reset_robots, reset_fire;
    double evac_reward, death_penalty, death_acc_penalty, alive_bonus;
    /* state */
    int *px, *py; double *health, *acc; uint8_t *saved, *dead;
    uint8_t* rmap;              /* [G] People.rmap as 0/1 */
    int robots[MAXR][2];        /* Map.robot_positions */
    int robot_position[2];      /* Map.robot_position (aliases robots[0] except right after reset, quirk Q7) */
    int fire_step, cur_step, prev_evac, prev_dead, episode, tick;
    /* scratch */
    int *grp_of_cell;           /* [G] -> group index or -1 */
    int *grp_cell, *grp_head, *grp_tail, *mv_next, *mv_dir;
    double* dist;
} orc_env;

#define CELL(e, x, y) ((x) * ((e)->W + 2) + (y))

/* map.py:85-92 */
static int check_valid(const orc_env* e, int x, int y) {
    if (x >= e->L + 1 || x <= 0 || y >= e->W + 1 || y <= 0) return 0;
    return !isinf(e->space[CELL(e, x, y)]);
}
/* map.py:93-113 */
static int check_savefy(const orc_env* e, int x, int y) {
    if (x >= e->L + 1) x = e->L + 1; else if (x <= 0) x = 0;
    if (y >= e->W + 1) y = e->W + 1; else if (y <= 0) y = 0;
    for (int k = 0; k < e->n_exits; ++k)
        if (abs(x - e->exits[2 * k]) <= 1 && abs(y - e->exits[2 * k + 1]) <= 1) return 1;
    return 0;
}
static double table_lookup(const int box[4], const double* tab, int n_steps, int step, int x, int y) {
    if (step >= n_steps) step = n_steps - 1;
    int rx = x - box[0], ry = y - box[1];
    if (rx < 0 || ry < 0 || rx >= box[2] || ry >= box[3]) return 0.0;
    return tab[((size_t)step * box[2] + rx) * box[3] + ry];
}

orc_env* orc_env_new(int L, int W, int N, int R, const double* space, const uint8_t* barrier, int n_exits,
                     const int32_t* exits, const int32_t* obs_exit, int n_fire_steps, const int32_t* ctr_box,
                     const double* danger_ctr, const int32_t* int_box, const double* danger_int,
                     const int32_t* robot_range, const int32_t* robot_start, const int32_t* reset_obs_center,
                     uint64_t seed, uint32_t env_id, int max_steps, int reset_robots, int reset_fire) {
    orc_env* e = (orc_env*)calloc(1, sizeof(orc_env));
    int G = (L + 2) * (W + 2);
    e->L = L; e->W = W; e->N = N; e->R = R;
    e->space = space; e->barrier = barrier; e->n_exits = n_exits; e->exits = exits;
    e->obs_exit[0] = obs_exit[0]; e->obs_exit[1] = obs_exit[1];
    e->n_fire_steps = n_fire_steps;
    memcpy(e->ctr_box, ctr_box, sizeof(e->ctr_box)); e->danger_ctr = danger_ctr;
    memcpy(e->int_box, int_box, sizeof(e->int_box)); e->danger_int = danger_int;
    e->robot_range[0] = robot_range[0]; e->robot_range[1] = robot_range[1];
    for (int r = 0; r < MAXR; ++r) { e->robot_start[r][0] = robot_start[2 * r]; e->robot_start[r][1] = robot_start[2 * r + 1]; }
    e->reset_obs_center[0] = reset_obs_center[0]; e->reset_obs_center[1] = reset_obs_center[1];
    e->seed = seed; e->env_id = env_id; e->max_steps = max_steps;
    e->reset_robots = reset_robots; e->reset_fire = reset_fire;
    e->evac_reward = 50.0; e->death_penalty = 200.0; e->death_acc_penalty = 0.5; e->alive_bonus = 1.0;   /* evacuation_env.py:16-19 */
    e->px = (int*)calloc(N, sizeof(int)); e->py = (int*)calloc(N, sizeof(int));
    e->health = (double*)calloc(N, sizeof(double)); e->acc = (double*)calloc(N, sizeof(double));
    e->saved = (uint8_t*)calloc(N, 1); e->dead = (uint8_t*)calloc(N, 1);
    e->rmap = (uint8_t*)calloc(G, 1);
    e->grp_of_cell = (int*)malloc(G * sizeof(int));
    for (int i = 0; i < G; ++i) e->grp_of_cell[i] = -1;
    e->grp_cell = (int*)calloc(N, sizeof(int)); e->grp_head = (int*)calloc(N, sizeof(int));
    e->grp_tail = (int*)calloc(N, sizeof(int)); e->mv_next = (int*)calloc(N, sizeof(int));
    e->mv_dir = (int*)calloc(N, sizeof(int));
    e->dist = (double*)calloc(N, sizeof(double));
    for (int r = 0; r < MAXR; ++r) { e->robots[r][0] = e->robot_start[r][0]; e->robots[r][1] = e->robot_start[r][1]; }
    e->robot_position[0] = e->robots[0][0]; e->robot_position[1] = e->robots[0][1];
    return e;
}
void orc_env_free(orc_env* e) {
    if (!e) return;
    free(e->px); free(e->py); free(e->health); free(e->acc); free(e->saved); free(e->dead); free(e->rmap);
    free(e->grp_of_cell); free(e->grp_cell); free(e->grp_head); free(e->grp_tail); free(e->mv_next); free(e->mv_dir);
    free(e->dist); free(e);
}
void orc_env_set_coefs(orc_env* e, double a, double b, double c, double d) {
    e->evac_reward = a; e->death_penalty = b; e->death_acc_penalty = c; e->alive_bonus = d;
}
void orc_env_set_robot(orc_env* e, int r, int x, int y) {
    e->robots[r][0] = x; e->robots[r][1] = y;
    if (r == 0) { e->robot_position[0] = x; e->robot_position[1] = y; }
}

/* evacuation_env.py:84-120 — one 11x11x6 window centred on (rx, ry), float64 */
static void get_state(const orc_env* e, int rx, int ry, double* out) {
    for (int i = 0; i < OBS_WIN; ++i)
        for (int j = 0; j < OBS_WIN; ++j) {
            int mx = rx + (i - 5), my = ry + (j - 5);
            double* o = out + (i * OBS_WIN + j) * OBS_CH;
            int valid = check_valid(e, mx, my);
            /* ch0: space/np.max(space) with max == inf -> 0.0 for every finite cell (quirk Q1) */
            o[0] = 0.0;
            o[1] = valid ? (double)e->rmap[CELL(e, mx, my)] : 0.0;
            o[2] = table_lookup(e->int_box, e->danger_int, e->n_fire_steps, e->fire_step, mx, my);
            int in_grid = (mx >= 0 && mx <= e->L + 1 && my >= 0 && my <= e->W + 1);
            o[3] = (!valid || (in_grid && e->barrier[CELL(e, mx, my)])) ? 1.0 : 0.0;
            o[4] = (mx == e->obs_exit[0] && my == e->obs_exit[1]) ? 1.0 : 0.0;
            o[5] = (i == 5 && j == 5) ? 1.0 : 0.0;
        }
}
/* evacuation_env_multi.py:44-53 — joint state; R == 1 is the single-robot env */
static void get_joint_state(const orc_env* e, double* out) {
    for (int r = 0; r < e->R; ++r) {
        int rx = (r == 0) ? e->robot_position[0] : e->robots[r][0];
        int ry = (r == 0) ? e->robot_position[1] : e->robots[r][1];
        get_state(e, rx, ry, out + (size_t)r * OBS_SIZE);
    }
}

/* evacuation_env.py:61-82 + people.py:158-194.  inject: [N][2] int16 cells or NULL */
void orc_env_reset(orc_env* e, const int16_t* inject, double* obs_out) {
    int G = (e->L + 2) * (e->W + 2);
    if (e->reset_robots) {                       /* evacuation_env_multi.py:35-36 */
        for (int r = 0; r < e->R; ++r) { e->robots[r][0] = e->robot_start[r][0]; e->robots[r][1] = e->robot_start[r][1]; }
        e->robot_position[0] = e->robots[0][0]; e->robot_position[1] = e->robots[0][1];
    } else {                                     /* evacuation_env.py:64 — robot_positions untouched (Q7) */
        e->robot_position[0] = e->reset_obs_center[0]; e->robot_position[1] = e->reset_obs_center[1];
    }
    if (e->reset_fire) e->fire_step = 0;         /* not in the reference (Q6) */
    memset(e->rmap, 0, G);
    for (int i = 0; i < e->N; ++i) {
        int x, y;
        if (inject) { x = inject[2 * i]; y = inject[2 * i + 1]; }
        else {
            for (int attempt = 0;; ++attempt) {   /* people.py:186-190 */
                uint32_t w[4];
                philox4x32(e->env_id, (uint32_t)e->episode, (uint32_t)i, STREAM_SPAWN + (attempt >> 1), e->seed, w);
                int j = 2 * (attempt & 1);
                x = 1 + (int)(((uint64_t)w[j] * (uint64_t)(e->L - 2)) >> 32);
                y = 1 + (int)(((uint64_t)w[j + 1] * (uint64_t)(e->W - 2)) >> 32);
                if (check_valid(e, x, y)) break;
            }
        }
        e->px[i] = x; e->py[i] = y;
        e->health[i] = 100.0; e->acc[i] = 0.0; e->saved[i] = 0; e->dead[i] = 0;
        e->rmap[CELL(e, x, y)] = 1;               /* duplicates allowed, flag semantics (Q2) */
    }
    e->cur_step = 0; e->prev_evac = 0; e->prev_dead = 0;
    e->episode += 1;
    if (obs_out) get_joint_state(e, obs_out);
}

/* map.py:160-202 */
static void move_robot(orc_env* e, int action, int rid) {
    if (action >= 0 && action <= 4) {
        int x = e->robots[rid][0], y = e->robots[rid][1];
        int nx = x, ny = y;
        if (action == 0) nx = x + 1; else if (action == 1) ny = y - 1;
        else if (action == 2) nx = x - 1; else if (action == 3) ny = y + 1;
        if (e->robot_range[0] <= nx && nx <= e->robot_range[1] && 0 <= ny && ny <= e->W && check_valid(e, nx, ny)) {
            e->robots[rid][0] = nx; e->robots[rid][1] = ny;
        }
    } else {
        return;                                   /* map.py:180-181: returns before the re-alias below */
    }
    if (rid == 0) { e->robot_position[0] = e->robots[0][0]; e->robot_position[1] = e->robots[0][1]; }
}

static const int MOVE_TO[8][2] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}, {1, -1}, {-1, -1}, {-1, 1}, {1, 1}};   /* map.py:11-19 */

/* people.py:255-297 */
static int find_best_direction(const orc_env* e, int person, int x, int y) {
    int best = -1;
    double max_score = -INFINITY;
    for (int dire = 0; dire < 8; ++dire) {
        int nx = x + MOVE_TO[dire][0], ny = y + MOVE_TO[dire][1];
        if (check_valid(e, nx, ny) && e->rmap[CELL(e, nx, ny)] == 0) {
            double delta_p = e->space[CELL(e, x, y)] - e->space[CELL(e, nx, ny)];
            double robot_effect = 0.0;
            double dist = INFINITY;
            for (int r = 0; r < e->R; ++r) {
                double dx = (double)(nx - e->robots[r][0]), dy = (double)(ny - e->robots[r][1]);
                double d = sqrt(dx * dx + dy * dy);
                if (d < dist) dist = d;
            }
            if (dist < 5.0) robot_effect = -20.0 / (dist + 0.1);          /* people.py:94-95,282-284 */
            uint32_t w[4];
            philox4x32(e->env_id, (uint32_t)e->tick, (uint32_t)person, (uint32_t)(dire >> 1), e->seed, w);
            double u = u53(w[2 * (dire & 1)], w[2 * (dire & 1) + 1]);
            double noise = -0.1 + (0.1 - -0.1) * u;                       /* random.uniform(-0.1, 0.1) */
            double score = delta_p * 5.0 + robot_effect + noise;
            if (score > max_score) { max_score = score; best = dire; }
        }
    }
    return best;
}

/* numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum) as used by
 * np.mean at evacuation_env.py:228 — third-party code not under /root/reference (numpy 2.3.5 here). */
static double pairwise_sum(const double* a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}
double orc_pairwise_sum(const double* a, long n) { return pairwise_sum(a, n); }

/* evacuation_env.py:174-288 */
static double calculate_reward(orc_env* e) {
    double reward = 0;
    const int N = e->N;
    int rx = e->robot_position[0], ry = e->robot_position[1];
    int cur_evac = 0, cur_dead = 0;
    for (int i = 0; i < N; ++i) { cur_evac += e->saved[i]; cur_dead += e->dead[i]; }
    int remaining = N - cur_evac - cur_dead;
    int new_evac = cur_evac - e->prev_evac;
    reward += new_evac * e->evac_reward;

    double guidance = 0;
    long n_rem = 0;
    for (int i = 0; i < N; ++i) {
        if (e->saved[i] || e->dead[i]) continue;
        double pxx = e->px[i] + 0.5, pyy = e->py[i] + 0.5;
        double dxr = pxx - rx, dyr = pyy - ry;
        double dist_to_robot = sqrt(dxr * dxr + dyr * dyr);
        if (dist_to_robot <= 5) {
            double dxe = pxx - e->obs_exit[0], dye = pyy - e->obs_exit[1];
            double dist_to_exit = sqrt(dxe * dxe + dye * dye);
            if (dist_to_exit > 20) guidance += 2.0;
            else if (dist_to_exit > 10) guidance += 1.5;
            else guidance += 1.0;
            if (e->health[i] < 80) guidance += 1.0;
            else if (e->health[i] < 60) guidance += 2.0;     /* unreachable (Q11) */
        }
        e->dist[n_rem++] = dist_to_robot;                     /* :225-229 uses robot_pos - pos: same squares */
    }
    reward += guidance;

    if (remaining > 0 && n_rem > 0) {
        double avg = pairwise_sum(e->dist, n_rem) / (double)n_rem;
        double dr = 2.0 - fabs(avg - 8.0) * 0.2;
        if (!(dr > 0)) dr = 0;                                 /* max(0, dr) */
        reward += dr;
    }
    if (remaining > 0) {
        double urgency = (double)remaining / (double)N;
        double time_penalty = -0.05 - (urgency * 0.1);
        reward += time_penalty;
    } else {
        reward -= 0.02;
    }
    double total_health = 0;                                   /* sum(): naive left-to-right */
    for (int i = 0; i < N; ++i) if (!e->dead[i]) total_health += e->health[i];
    if (N - cur_dead > 0) {
        double avg_health = total_health / (double)(N - cur_dead);
        reward += (avg_health - 90) * 0.05;
    }
    if (cur_evac == N) {
        double completion = 100;
        int rem_steps = 300 - e->cur_step; if (rem_steps < 0) rem_steps = 0;
        double time_bonus = rem_steps * 0.2;
        if (N > 0) {
            double all_health = 0;
            for (int i = 0; i < N; ++i) all_health += e->health[i];
            double final_avg = all_health / (double)N;
            double health_bonus = (final_avg - 80) * 1.0;
            reward += completion + time_bonus + health_bonus;
        } else reward += completion + time_bonus;
    }
    int new_deaths = cur_dead - e->prev_dead;
    reward -= new_deaths * e->death_penalty;
    reward -= cur_dead * e->death_acc_penalty;
    int survivors = N - cur_dead;
    reward += survivors * e->alive_bonus;
    if (e->cur_step > 0) {
        double eff = (double)cur_evac / (double)e->cur_step;
        if (eff > 0.1) reward += eff * 5;
    }
    e->prev_evac = cur_evac; e->prev_dead = cur_dead;
    return reward;
}

/* evacuation_env.py:122-172 / evacuation_env_multi.py:55-89 */
void orc_env_step(orc_env* e, const int32_t* actions, double* obs_out, double* reward_out, uint8_t* done_out) {
    const int N = e->N;
    for (int r = 0; r < e->R; ++r) move_robot(e, actions[r], r);

    /* people.py:203-207 — phase 1: health + speed */
    double* speed = e->dist;   /* scratch reuse: Person.speed of this step */
    for (int i = 0; i < N; ++i) {
        if (e->saved[i] || e->dead[i]) continue;
        double danger = table_lookup(e->ctr_box, e->danger_ctr, e->n_fire_steps, e->fire_step, e->px[i], e->py[i]);
        double h = e->health[i];
        if (danger > 0) {                                            /* people.py:61-88 */
            uint32_t w[4];
            philox4x32(e->env_id, (uint32_t)e->tick, (uint32_t)i, STREAM_HEALTH, e->seed, w);
            double u = u53(w[0], w[1]);
            double loss;
            if (danger >= 0.8) loss = danger * 50.0 + (1.0 + (3.0 - 1.0) * u);
            else if (danger >= 0.5) loss = danger * 40.0 + (0.8 + (2.0 - 0.8) * u);
            else if (danger >= 0.2) loss = danger * 30.0 + (0.5 + (1.5 - 0.5) * u);
            else loss = danger * 20.0 + (0.2 + (1.0 - 0.2) * u);
            if (h < 50) loss *= 1.2; else if (h < 25) loss *= 1.4;
            h -= loss;
            if (h <= 0) { h = 0; e->dead[i] = 1; } else if (h <= 8.0) e->dead[i] = 1;
            h = fmax(0, fmin(h, 100));
            e->health[i] = h;
        }
        if (e->dead[i]) continue;
        if (h < 20) speed[i] = 0.4;                                   /* people.py:38-44 */
        else { double f = 0.3 + 0.7 * (h / 100.0); speed[i] = 1.0 * f; }
    }
    /* people.py:210-230 — phase 2: proposals, move_plan in insertion order */
    int n_grp = 0;
    for (int i = 0; i < N; ++i) {
        if (e->saved[i] || e->dead[i]) continue;
        e->acc[i] += speed[i] * 0.5;
        if (e->acc[i] >= 1.0) {
            e->acc[i] -= 1.0;
            int d = find_best_direction(e, i, e->px[i], e->py[i]);
            if (d >= 0) {
                int c = CELL(e, e->px[i] + MOVE_TO[d][0], e->py[i] + MOVE_TO[d][1]);
                int g = e->grp_of_cell[c];
                if (g < 0) { g = n_grp++; e->grp_of_cell[c] = g; e->grp_cell[g] = c; e->grp_head[g] = i; }
                else e->mv_next[e->grp_tail[g]] = i;
                e->grp_tail[g] = i; e->mv_next[i] = -1; e->mv_dir[i] = d;
            }
        }
    }
    /* people.py:238-249 — phase 4: shuffle each group, movers[0] moves (execute_move :299-314) */
    for (int g = 0; g < n_grp; ++g) {
        int c = e->grp_cell[g];
        e->grp_of_cell[c] = -1;
        int win = -1; uint32_t best = 0;
        for (int i = e->grp_head[g]; i >= 0; i = e->mv_next[i]) {
            uint32_t w[4];
            philox4x32(e->env_id, (uint32_t)e->tick, (uint32_t)i, STREAM_HEALTH, e->seed, w);
            if (win < 0 || w[2] < best) { win = i; best = w[2]; }     /* keyed shuffle law: min priority, ties -> lower index */
        }
        int nx = c / (e->W + 2), ny = c % (e->W + 2);
        e->rmap[CELL(e, e->px[win], e->py[win])] = 0;
        e->rmap[c] = 1;
        e->px[win] = nx; e->py[win] = ny;
        if (check_savefy(e, nx, ny)) { e->saved[win] = 1; e->rmap[c] = 0; }
    }
    /* evacuation_env.py:138-142 — both fire models step (fire_model.py:63-67) */
    if (e->fire_step < e->n_fire_steps - 1) e->fire_step += 1;

    double reward = calculate_reward(e);
    e->cur_step += 1;
    e->tick += 1;
    int cur_evac = 0, cur_dead = 0;
    for (int i = 0; i < N; ++i) { cur_evac += e->saved[i]; cur_dead += e->dead[i]; }
    int done = (cur_evac + cur_dead == N) || (e->cur_step >= e->max_steps);   /* time = 0.5*step >= 600 */
    if (reward_out) *reward_out = reward;
    if (done_out) *done_out = (uint8_t)done;
    if (obs_out) get_joint_state(e, obs_out);
}

/* ---- state export ------------------------------------------------------- */
void orc_env_export(const orc_env* e, int16_t* px, int16_t* py, double* health, double* acc, uint8_t* flags,
                    uint8_t* rmap, int32_t* robots, int32_t* scalars) {
    for (int i = 0; i < e->N; ++i) {
        px[i] = (int16_t)e->px[i]; py[i] = (int16_t)e->py[i];
        health[i] = e->health[i]; acc[i] = e->acc[i];
        flags[i] = (uint8_t)(e->saved[i] | (e->dead[i] << 1));
    }
    memcpy(rmap, e->rmap, (size_t)(e->L + 2) * (e->W + 2));
    for (int r = 0; r < MAXR; ++r) { robots[2 * r] = e->robots[r][0]; robots[2 * r + 1] = e->robots[r][1]; }
    int ev = 0, dd = 0;
    for (int i = 0; i < e->N; ++i) { ev += e->saved[i]; dd += e->dead[i]; }
    scalars[0] = e->fire_step; scalars[1] = e->cur_step; scalars[2] = e->prev_evac; scalars[3] = e->prev_dead;
    scalars[4] = e->episode; scalars[5] = e->tick; scalars[6] = ev; scalars[7] = dd;
}
void orc_env_set_fire_step(orc_env* e, int s) { e->fire_step = s; }

/* ---- batch helpers (bench cpu_baseline: all host threads, pthreads) -------- */
#include <pthread.h>
typedef struct { orc_env** envs; int n; const int32_t* actions; int auto_reset; float* obs32; double* reward;
                 uint8_t* done; int do_reset; volatile int* next; } orc_job;
static void* orc_worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    for (;;) {
        int k0 = __sync_fetch_and_add(j->next, 4);
        if (k0 >= j->n) break;
        int k1 = k0 + 4 < j->n ? k0 + 4 : j->n;
        for (int k = k0; k < k1; ++k) {
            orc_env* e = j->envs[k];
            double obs[MAXR * OBS_SIZE];
            if (j->do_reset) orc_env_reset(e, NULL, obs);
            else {
                double r; uint8_t d;
                orc_env_step(e, j->actions + (size_t)k * e->R, obs, &r, &d);
                if (d && j->auto_reset) orc_env_reset(e, NULL, obs);
                if (j->reward) j->reward[k] = r;
                if (j->done) j->done[k] = d;
            }
            if (j->obs32) for (int q = 0; q < e->R * OBS_SIZE; ++q) j->obs32[(size_t)k * e->R * OBS_SIZE + q] = (float)obs[q];
        }
    }
    return NULL;
}
static void orc_run(orc_job* j, int n_threads) {
    volatile int next = 0;
    j->next = &next;
    if (n_threads <= 1) { orc_worker(j); return; }
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, orc_worker, j);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
}
/* Steps n envs once on n_threads host threads; env k reads actions[k*R..].  auto_reset re-spawns
 * finished envs like the batched GPU path.  obs32 / reward / done may be NULL. */
void orc_batch_step(orc_env** envs, int n, const int32_t* actions, int auto_reset, float* obs32, double* reward,
                    uint8_t* done, int n_threads) {
    orc_job j = {envs, n, actions, auto_reset, obs32, reward, done, 0, NULL};
    orc_run(&j, n_threads);
}
void orc_batch_reset(orc_env** envs, int n, float* obs32, int n_threads) {
    orc_job j = {envs, n, NULL, 0, obs32, NULL, NULL, 1, NULL};
    orc_run(&j, n_threads);
}
