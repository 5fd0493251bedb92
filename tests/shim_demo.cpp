// shim_demo.cpp -- exercises cognn_b200/host/shim/cognn_shim.h the way the reference's operators use the primitive API:
// an ALICE thread (graph owner, tile 0) and a BOB thread (helper, tile 1) issue the same call sequence as
// ssk.h:736-821 + optimize-gcn/gcn.h:198-494 for one forward GAS iteration over the owner's local edges, with the
// reference's in/out aliasing.  Checks: (1) the linear part (expand -> copy -> OGA -> extract -> conditional add) is
// exact mod 2^64 on reconstructed values, (2) the fixed-point results track a float64 computation, (3) the fused
// engine kernel cgb_gather_sum gives the same reconstructed update.  Built and run by tests/test_gpu_shim.py.
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <random>
#include <thread>

#define COGNN_SHIM_IDEAL_NONLINEAR  // test data only: the 2PC-residual stand-ins are NOT secure
#include "../cognn_b200/host/shim/cognn_shim.h"

struct Party {
    ShareVecVec X, W, Xp, upd, out;
    DoubleTensor plain;
    ShareVecVec relu, p, pmy, masked, g_empty;  // outputs of the three stand-ins
};

static void run_party(cognn_shim::Runtime* rt, int party, uint64_t coTid, const std::vector<uint64_t>& localVertexPos,
                      const std::vector<uint64_t>& updateSrcVertexPos, const std::vector<uint64_t>& updateDstVertexPos,
                      const std::vector<bool>& isGatherDstVertexDummy, const std::vector<uint64_t>& inDeg, Party* st) {
    cognn_shim::Runtime::bind_thread(rt);
    const bool alice = party == sci::ALICE;
    const uint32_t D = (uint32_t)st->W[0].size();
    const size_t E = updateSrcVertexPos.size(), n = localVertexPos.size();
    ShareVecVec& v = st->X;
    // PreScatterComp (gcn.h:233-239): output aliases the input, and the width changes from F to H
    sci::twoPartyGCNMatMul(v, st->W, v, coTid, party);
    st->Xp = v;
    // Scatter preparation (ssk.h:752): rows N_p -> E_local
    ShareVecVec updateSrc, dup, localUpdate;
    if (alice) client_oblivious_mapper_online(localVertexPos, updateSrcVertexPos, v, updateSrc, D, 0, 0, coTid);
    else server_oblivious_mapper_online(v, updateSrc, 0, 0, coTid);
    dup = updateSrc;  // ScatterComp of CoGNN-Opt is a copy (gcn.h:300)
    std::vector<uint64_t> zeroPos(E, 0);
    dup = prefix_network_aggregate(alice ? updateDstVertexPos : zeroPos, dup, AggregationOp::ADD_AGG, coTid, party, true);
    // premerged extraction (ssk.h:818): rows E_local -> N_p
    if (alice) client_oblivious_mapper_online(updateDstVertexPos, localVertexPos, dup, localUpdate, D, 0, 2, coTid);
    else server_oblivious_mapper_online(dup, localUpdate, 0, 2, coTid);
    st->upd = localUpdate;
    // GatherComp (gcn.h:454-483)
    std::vector<bool> cond(n, true);
    if (alice)
        for (size_t i = 0; i < n; ++i) cond[i] = !isGatherDstVertexDummy[i];
    sci::twoPartyGCNCondVectorAddition(v, localUpdate, cond, v, coTid, party);
    std::vector<uint64_t> normalizer(n, 0);
    if (alice)
        for (size_t i = 0; i < n; ++i)
            normalizer[i] = inDeg[i] == 0 ? 0 : CryptoUtil::encodeDoubleAsFixedPoint(pow((double)inDeg[i] + 1, -0.5));
    sci::twoPartyGCNVectorScale(v, normalizer, v, true, coTid, party);
    st->out = v;
    sci::getPlainShareVecVec(v, st->plain, coTid, party);
    // the three 2PC-residual calls of ApplyComp (gcn.h:549, 578/591, 705), here as ideal-functionality stand-ins
    sci::twoPartyGCNRelu(v, st->relu, coTid, party);
    ShareVecVec label(n, ShareVec(D, 0));
    if (alice)
        for (size_t i = 0; i < n; ++i) label[i] = toShareVec((int)(i % D), (int)D);  // ALICE: one-hot, BOB: zeros
    sci::twoPartyGCNForwardNNPredictionWithoutWeight(v, label, st->p, st->pmy, coTid, party);
    ShareTensor noWeight;
    sci::twoPartyGCNBackwardNNWithoutAH(st->pmy, v, noWeight, st->masked, st->g_empty, /*isFirstLayer=*/true, coTid, party);
}

int main() {
    if (cgb_device_count() == 0) {
        printf("no CUDA device\n");
        return 2;
    }
    const uint32_t key[8] = {45, 0, 0, 0, 0, 0, 0, 7};
    std::mt19937_64 rng(123);
    const size_t n = 300, F = 24, H = 16, E = 2000;
    // owner-local index vectors: edges grouped by destination ascending; vertex ids = 2 * row (party 0 of vid % 2)
    std::vector<uint64_t> localVertexPos(n), src(E), dst(E), inDeg(n, 0);
    for (size_t i = 0; i < n; ++i) localVertexPos[i] = 2 * i;
    std::vector<std::pair<uint64_t, uint64_t>> edges(E);
    for (auto& e : edges) e = {2 * (rng() % n), 2 * ((rng() % n) * (rng() % 4 != 0) % n)};  // skewed destinations
    std::sort(edges.begin(), edges.end(), [](auto& a, auto& b) { return a.second != b.second ? a.second < b.second : a.first < b.first; });
    std::vector<bool> dummy(n, true);
    for (size_t e = 0; e < E; ++e) {
        src[e] = edges[e].first;
        dst[e] = edges[e].second;
        inDeg[dst[e] / 2]++;
        dummy[dst[e] / 2] = false;
    }
    // -r 1: a vertex without local in-edge gets a dummy self edge (ssk.h:412-418); it flows through expand and OGA and
    // is dropped at Gather by isGatherDstVertexDummy
    std::vector<uint64_t> usrc, udst;
    {
        size_t e = 0;
        for (size_t i = 0; i < n; ++i) {
            if (dummy[i]) {
                usrc.push_back(2 * i);
                udst.push_back(2 * i);
                inDeg[i] += 1;
            }
            while (e < E && dst[e] == 2 * i) {
                usrc.push_back(src[e]);
                udst.push_back(dst[e]);
                ++e;
            }
        }
    }
    // plaintext inputs, shared with CryptoUtil::intoShares on the owner's runtime
    cognn_shim::Runtime rtA(0, 2, 0, key), rtB(1, 2, 0, key);
    cognn_shim::InProcPipe pipe;
    rtA.connect(1, sci::ALICE, &pipe.a);
    rtB.connect(0, sci::BOB, &pipe.b);
    cognn_shim::Runtime::bind_thread(&rtA);
    std::uniform_real_distribution<double> ud(-1.0, 1.0);
    std::vector<std::vector<double>> Xd(n, std::vector<double>(F)), Wd(F, std::vector<double>(H));
    Party A, B;
    A.X.assign(n, ShareVec(F)); B.X.assign(n, ShareVec(F));
    A.W.assign(F, ShareVec(H)); B.W.assign(F, ShareVec(H));
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < F; ++j) {
            Xd[i][j] = ud(rng);
            CryptoUtil::intoShares(Xd[i][j], A.X[i][j], B.X[i][j]);
        }
    for (size_t i = 0; i < F; ++i)
        for (size_t j = 0; j < H; ++j) {
            Wd[i][j] = ud(rng) * 0.3;
            CryptoUtil::intoShares(Wd[i][j], A.W[i][j], B.W[i][j]);
        }
    std::thread ta(run_party, &rtA, sci::ALICE, 1, localVertexPos, usrc, udst, dummy, inDeg, &A);
    std::thread tb(run_party, &rtB, sci::BOB, 0, localVertexPos, usrc, udst, dummy, inDeg, &B);
    ta.join();
    tb.join();

    int bad = 0;
    // (1) exact linear part on reconstructed values: upd[v] = sum over REAL in-edges of Xp[src]; dummy rows carry Xp[v]
    std::vector<std::vector<uint64_t>> Xp(n, std::vector<uint64_t>(H)), want(n, std::vector<uint64_t>(H, 0));
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < H; ++j) Xp[i][j] = A.Xp[i][j] + B.Xp[i][j];
    for (size_t e = 0; e < usrc.size(); ++e)
        for (size_t j = 0; j < H; ++j) want[udst[e] / 2][j] += Xp[usrc[e] / 2][j];
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < H; ++j)
            if (A.upd[i][j] + B.upd[i][j] != want[i][j]) ++bad;
    printf("linear part (expand, OGA, extract) exact mismatches: %d\n", bad);
    // (2) float64 reference of the whole step
    double max_err = 0;
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < H; ++j) {
            auto xw = [&](size_t r) {
                double s = 0;
                for (size_t k = 0; k < F; ++k) s += Xd[r][k] * Wd[k][j];
                return s;
            };
            double v = xw(i);
            if (!dummy[i])
                for (size_t e = 0; e < E; ++e)
                    if (dst[e] == 2 * i) v += xw(src[e] / 2);
            v *= pow((double)inDeg[i] + 1, -0.5);
            max_err = std::max(max_err, std::fabs(v - A.plain[i][j]));
            const double merged = CryptoUtil::mergeShareAsDouble(A.out[i][j], B.out[i][j]);
            if (merged != A.plain[i][j]) ++bad;
        }
    }
    printf("max |secure - float64| = %g\n", max_err);
    if (max_err > 5e-2) ++bad;
    if (!B.plain.empty()) ++bad;  // only ALICE learns the opened values (gcn.h:604-605)
    // (3) the fused kernel of the engine gives the same reconstructed update from the same reconstructed Xp
    {
        cgb_ctx* c = nullptr;
        cgb_ctx_create(0, &c);
        std::vector<uint32_t> rowptr(n + 1, 0), col;
        for (size_t e = 0; e < E; ++e) rowptr[dst[e] / 2 + 1]++;
        for (size_t i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
        for (size_t e = 0; e < E; ++e) col.push_back((uint32_t)(src[e] / 2));
        cgb_csr* csr = nullptr;
        if (cgb_csr_create(c, rowptr.data(), col.data(), n, E, n, &csr) != CGB_OK) ++bad;
        std::vector<uint64_t> flat(n * H), y(n * H);
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < H; ++j) flat[i * H + j] = Xp[i][j];
        if (cgb_host_gather_sum(c, csr, flat.data(), nullptr, y.data(), H) != CGB_OK) ++bad;
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < H; ++j)
                if (!dummy[i] && y[i * H + j] != want[i][j]) ++bad;
        cgb_csr_destroy(c, csr);
        cgb_ctx_destroy(c);
    }
    // (4) the stand-ins on reconstructed values: ReLU and the ReLU' mask are exact, the softmax rows sum to one and match
    //     a double-precision softmax of the opened logits within the fixed-point resolution
    {
        int bad_relu = 0, bad_mask = 0;
        double max_p_err = 0;
        const double sc = (double)(1ull << SCALER_BIT_LENGTH);
        for (size_t i = 0; i < n; ++i) {
            double mx = -1e300, tot = 0;
            for (size_t j = 0; j < H; ++j) mx = std::max(mx, (double)(int64_t)(A.out[i][j] + B.out[i][j]) / sc);
            for (size_t j = 0; j < H; ++j) tot += std::exp((double)(int64_t)(A.out[i][j] + B.out[i][j]) / sc - mx);
            for (size_t j = 0; j < H; ++j) {
                const uint64_t vv = A.out[i][j] + B.out[i][j];
                const uint64_t want_relu = (int64_t)vv > 0 ? vv : 0;
                if (A.relu[i][j] + B.relu[i][j] != want_relu) ++bad_relu;
                const uint64_t pj = A.p[i][j] + B.p[i][j], dj = A.pmy[i][j] + B.pmy[i][j];
                max_p_err = std::max(max_p_err, std::fabs((double)(int64_t)pj / sc - std::exp((double)(int64_t)vv / sc - mx) / tot));
                if (dj != pj - ((i % H) == j ? (1ull << SCALER_BIT_LENGTH) : 0ull)) ++bad_mask;
                const uint64_t want_m = (int64_t)vv > 0 ? dj : 0;
                if (A.masked[i][j] + B.masked[i][j] != want_m) ++bad_mask;
            }
        }
        printf("stand-ins: relu mismatches %d, p - y / mask mismatches %d, max |p - softmax| = %g\n", bad_relu, bad_mask, max_p_err);
        if (bad_relu || bad_mask || max_p_err > 2.0 / sc || !A.g_empty.empty()) ++bad;
    }
    printf(bad == 0 ? "SHIM_DEMO_OK\n" : "SHIM_DEMO_FAILED (%d)\n", bad);
    return bad == 0 ? 0 : 1;
}
