// Static floor field — host, init time.  Replaces Map.Init_Potential of the reference
// (Louvre_Evacuation/envs/map.py:127-148): 8-connected Dijkstra seeded with 1 at every exit,
// step cost 1.0 (axis moves) / 1.4 (diagonals) accumulated in float64 exactly as
// `current_dist + cost` does, then += add_term on every reached cell.
//
// The reference pushes duplicates and never skips stale heap entries; because float64 addition is
// monotone (a <= b  =>  fl(a+c) <= fl(b+c)) the fixed point min_u fl(d[u] + c(u,v)) is unique, so a
// label-setting search with stale-entry skipping returns bit-identical distances.
#include <cmath>
#include <limits>
#include <queue>
#include <tuple>
#include <vector>
#include "common.h"

extern "C" int mq_floor_field(int32_t L, int32_t W, const uint8_t* wall, const int32_t* exits, int32_t n_exits,
                              const double* add_term, double* space_out) {
    MQ_REQUIRE(L > 0 && W > 0 && wall && exits && n_exits > 0 && space_out, "mq_floor_field: bad argument");
    const int stride = W + 2;
    const int G = (L + 2) * stride;
    const double inf = std::numeric_limits<double>::infinity();
    for (int i = 0; i < G; ++i) space_out[i] = inf;
    // map.py:11-19 MoveTO order; cost 1.0 for the first four, 1.4 for the diagonals (map.py:137)
    static const int mv[8][2] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}, {1, -1}, {-1, -1}, {-1, 1}, {1, 1}};
    using Item = std::tuple<double, int, int>;
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
    for (int k = 0; k < n_exits; ++k) {
        int ex = exits[2 * k], ey = exits[2 * k + 1];
        MQ_REQUIRE(ex >= 0 && ex <= L + 1 && ey >= 0 && ey <= W + 1, "mq_floor_field: exit %d outside the grid", k);
        space_out[ex * stride + ey] = 1.0;   // map.py:131
        heap.emplace(1.0, ex, ey);
    }
    while (!heap.empty()) {
        auto [d, x, y] = heap.top();
        heap.pop();
        if (d > space_out[x * stride + y]) continue;
        for (int i = 0; i < 8; ++i) {
            int nx = x + mv[i][0], ny = y + mv[i][1];
            // Map.Check_Valid on the pre-search grid (map.py:85-92): inside 1..L x 1..W and not a wall
            if (nx >= L + 1 || nx <= 0 || ny >= W + 1 || ny <= 0) continue;
            if (wall[nx * stride + ny]) continue;
            double nd = d + (i < 4 ? 1.0 : 1.4);
            if (nd < space_out[nx * stride + ny]) {
                space_out[nx * stride + ny] = nd;
                heap.emplace(nd, nx, ny);
            }
        }
    }
    if (add_term)
        for (int i = 0; i < G; ++i)
            if (space_out[i] != inf) space_out[i] += add_term[i];   // map.py:145-147
    return MQ_OK;
}
