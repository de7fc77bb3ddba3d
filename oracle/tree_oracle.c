/* ORACLE (test infrastructure, never shipped or measured as the product).
 *
 * Plain-C restatement of the reference's tree search with injected network outputs
 * ("teacher forcing"): MCTS.run's selection / expansion / backup (self_play.py:319-355),
 * ucb_score (self_play.py:380-404), backpropagate (self_play.py:406-430), MinMaxStats
 * (self_play.py:553-570), root noise mixing (self_play.py:476).  IEEE fp64, one rounding per
 * Python operation - compile with -ffp-contract=off (oracle/build_c.py does).
 *
 * Exists so full-size batches (4096 games x 50 simulations, 1024 x 200) can be checked
 * against the device in milliseconds.  Pinned against oracle/mcts.py, which is pinned against
 * fixtures produced by the reference itself (tests/test_tree_oracle_c.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static int tie_index(uint64_t seed, int64_t game, int move, int sim, int depth, int n) {
    uint32_t c[4] = {(uint32_t)game, (uint32_t)move, (uint32_t)sim, (uint32_t)depth};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x7169E001u);
    return (int)(((uint64_t)c[0] * (uint32_t)n) >> 32);
}

typedef struct {
    int visit, to_play, kids, nkids;    /* kids = index of first child, -1 if not expanded */
    double vsum, prior, reward;
    int action;
} Node;

int mz_oracle_tree_search(
    int n, int N, int A, int P, double discount, double pb_c_base, double pb_c_init, double frac,
    const unsigned char* legal, const int* to_play, const double* noise, const int* first_index,
    uint64_t seed, const int64_t* game_id, const int* move_index,
    const float* t_root_reward, const float* t_root_priors, const float* t_value, const float* t_reward,
    const float* t_priors,
    int* visit_counts, double* root_value, int* max_depth, int* ties, double* range,
    int* path_depth, unsigned char* path_actions, int D)
{
    Node* pool = (Node*)malloc(sizeof(Node) * (size_t)(1 + (N + 1) * A));
    int* path = (int*)malloc(sizeof(int) * (size_t)(N + 2));
    double* score = (double*)malloc(sizeof(double) * (size_t)A);
    int* tied = (int*)malloc(sizeof(int) * (size_t)A);
    if (!pool || !path || !score || !tied) return -1;
    for (int g = 0; g < n; ++g) {
        int used = 1, deepest = 0, nties = 0;
        double lo = INFINITY, hi = -INFINITY;
        Node* root = &pool[0];
        memset(root, 0, sizeof(Node));
        root->to_play = to_play ? to_play[g] : 0;
        root->reward = t_root_reward[g];
        root->kids = used;
        for (int a = 0; a < A; ++a) {
            if (legal && !legal[(size_t)g * A + a]) continue;
            Node* c = &pool[used++];
            memset(c, 0, sizeof(Node));
            c->kids = -1; c->to_play = -1; c->action = a;
            c->prior = (double)t_root_priors[(size_t)g * A + a];
            if (noise) c->prior = c->prior * (1 - frac) + noise[(size_t)g * A + a] * frac;
        }
        root->nkids = used - root->kids;
        const int64_t gid = game_id ? game_id[g] : g;
        const int mv = move_index ? move_index[g] : 0;
        for (int sim = 0; sim < N; ++sim) {
            int vtp = root->to_play, depth = 0, cur = 0;
            path[0] = 0;
            while (pool[cur].kids >= 0) {
                Node* p = &pool[cur];
                double best = -INFINITY;
                for (int k = 0; k < p->nkids; ++k) {
                    Node* c = &pool[p->kids + k];
                    double pbc = log((p->visit + pb_c_base + 1) / pb_c_base) + pb_c_init;
                    pbc *= sqrt((double)p->visit) / (c->visit + 1);
                    double s = pbc * c->prior;
                    if (c->visit > 0) {
                        double q = c->vsum / c->visit;
                        double v = c->reward + discount * (P == 1 ? q : -q);
                        if (hi > lo) v = (v - lo) / (hi - lo);
                        s = s + v;
                    } else {
                        s = s + 0;
                    }
                    score[k] = s;
                    if (s > best) best = s;
                }
                int nt = 0;
                for (int k = 0; k < p->nkids; ++k) if (score[k] == best) tied[nt++] = k;
                int pick;
                if (nt == 1) pick = tied[0];
                else if (sim == 0 && depth == 0 && first_index && first_index[g] >= 0)
                    pick = tied[first_index[g] < nt ? first_index[g] : nt - 1];
                else {
                    pick = tied[tie_index(seed, gid, mv, sim, depth, nt)];
                    if (!(sim == 0 && depth == 0)) nties++;
                }
                cur = p->kids + pick;
                if (path_actions && depth < D) path_actions[((size_t)g * N + sim) * D + depth] = (unsigned char)pool[cur].action;
                depth++;
                path[depth] = cur;
                vtp = (vtp + 1 < P) ? vtp + 1 : 0;
            }
            /* expand with the injected outputs of this simulation */
            Node* leaf = &pool[cur];
            leaf->to_play = vtp;
            leaf->reward = (double)t_reward[(size_t)g * N + sim];
            leaf->kids = used;
            leaf->nkids = A;
            for (int a = 0; a < A; ++a) {
                Node* c = &pool[used++];
                memset(c, 0, sizeof(Node));
                c->kids = -1; c->to_play = -1; c->action = a;
                c->prior = (double)t_priors[((size_t)g * N + sim) * A + a];
            }
            /* backup */
            double value = (double)t_value[(size_t)g * N + sim];
            for (int j = depth; j >= 0; --j) {
                Node* nd = &pool[path[j]];
                if (P == 1) {
                    nd->vsum += value;
                    nd->visit += 1;
                    double m = nd->reward + discount * (nd->vsum / nd->visit);
                    if (m > hi) hi = m;
                    if (m < lo) lo = m;
                    value = nd->reward + discount * value;
                } else {
                    const int same = nd->to_play == vtp;
                    nd->vsum += same ? value : -value;
                    nd->visit += 1;
                    double m = nd->reward + discount * -(nd->vsum / nd->visit);
                    if (m > hi) hi = m;
                    if (m < lo) lo = m;
                    value = (same ? -nd->reward : nd->reward) + discount * value;
                }
            }
            if (depth > deepest) deepest = depth;
            if (path_depth) path_depth[(size_t)g * N + sim] = depth;
        }
        for (int a = 0; a < A; ++a) visit_counts[(size_t)g * A + a] = 0;
        for (int k = 0; k < root->nkids; ++k) visit_counts[(size_t)g * A + pool[root->kids + k].action] = pool[root->kids + k].visit;
        root_value[g] = root->visit ? root->vsum / root->visit : 0.0;
        max_depth[g] = deepest;
        ties[g] = nties;
        range[2 * g] = lo; range[2 * g + 1] = hi;
    }
    free(pool); free(path); free(score); free(tied);
    return 0;
}
