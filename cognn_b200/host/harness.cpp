// harness.cpp -- command-line driver with the reference's interface: the counterpart of
// algo_kernels/common_harness/harness.cpp:50-212 + include/harness.h:91-220 + include/engine.h:143-222 of the reference.
//
//   gcn-optimize-b200 -t T -g T -i me -m iters -p 1 -s <setting> [-n 1] [-c 1] -r 1 [-u] edge vertex part result config
//
// Same flags, same positional files, same input formats (graph_io_util.h:66-147: `.edge` = "src dst [w]", `.part` = "vid
// tid", '#' comments; kernel_harness.h:37-44: `.vertex` = "vid f_0 .. f_{F-1} label"; task.h:106-169: "key : value"
// config) and the reference's log lines ("::iteration took", accuracy lines), so tools/plot/*.py keep working.
// One process per party (as tools/tmp_run_cluster.py:105-151 launches them), one GPU per process, NCCL between them: the
// 128-byte NCCL id is published by party 0 in a file next to the result file.  COGNN_B200_PLANE=loopback runs all T parties
// in this one process on one GPU instead (then -i is ignored).  -r 1 is required (the only dummy-edge mode every eval
// script uses); -n / -c are accepted and ignored (no OM preprocessing files, no WAN emulation).
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"

using namespace cognn;

static std::istream& nextEffectiveLine(std::istream& input, std::string& line) {  // graph_io_util.h:17-22
    do {
        std::getline(input, line);
    } while (input && (line.empty() || line.at(0) == '#'));
    return input;
}

static void printHelp() {
    std::cerr << "Usage: ./gcn-optimize-b200 -t <threadCount> -g <graphTileCount> [options] <edgelistFile> <vertexlistFile> "
                 "[partitionFile] [outputFile] [GNNConfigFile]\n"
                 "Options:\n\t-t <threadCount>\tNumber of parties.\n\t-g <graphTileCount>\tNumber of graph tiles.\n"
                 "\t-i <tileIndex>\tIndex of this party.\n\t-m <maxIter>\tMaximum iteration number.\n\t-p <numParts>\n"
                 "\t-s <setting>\n\t-n <0|1>\tno preprocess (ignored)\n\t-c <0|1>\tcluster (ignored)\n\t-r <0|1>\tno dummy edges (must be 1)\n"
                 "\t-u\tTreat as undirected graph.\n\t-h\tPrint this help message.\n";
}

int main(int argc, char* argv[]) {
    size_t threadCount = 0, graphTileCount = 0, tileIndex = 0;
    uint64_t maxIters = 1000;
    uint32_t numParts = 16;
    bool undirected = false, noPreprocess = false, isCluster = false, isNoDummyEdge = false;
    std::string setting;
    int ch;
    opterr = 0;
    while ((ch = getopt(argc, argv, "t:g:i:m:p:s:n:c:r:uh")) != -1) {
        uint32_t flag = 0;
        switch (ch) {
            case 't': std::stringstream(optarg) >> threadCount; break;
            case 'g': std::stringstream(optarg) >> graphTileCount; break;
            case 'i': std::stringstream(optarg) >> tileIndex; break;
            case 'm': std::stringstream(optarg) >> maxIters; break;
            case 'p': std::stringstream(optarg) >> numParts; break;
            case 's':
                std::stringstream(optarg) >> setting;
                /* falls through, as in the reference (harness.h:140-146) */
            case 'n':
                std::stringstream(optarg) >> flag;
                if (flag == 1) noPreprocess = true;
                break;
            case 'c':
                std::stringstream(optarg) >> flag;
                if (flag == 1) isCluster = true;
                break;
            case 'r':
                std::stringstream(optarg) >> flag;
                if (flag == 1) isNoDummyEdge = true;
                break;
            case 'u': undirected = true; break;
            case 'h':
            default: printHelp(); return -1;
        }
    }
    (void)noPreprocess; (void)isCluster; (void)numParts;
    if (threadCount == 0 || graphTileCount == 0) {
        std::cerr << "Must specify number of threads and number of graph tiles." << std::endl;
        printHelp();
        return -1;
    }
    if (graphTileCount % threadCount != 0) {
        std::cerr << "Number of threads must be a divisor of number of graph tiles." << std::endl;
        printHelp();
        return -1;
    }
    argc -= optind;
    argv += optind;
    if (argc < 1) {
        std::cerr << "Must specify an input edge list file." << std::endl;
        printHelp();
        return -1;
    }
    if (argc < 5) {
        std::cerr << "Must specify edge, vertex, partition, result and GNN config files." << std::endl;
        return -1;
    }
    const std::string edgelistFile = argv[0], vertexlistFile = argv[1], partitionFile = argv[2], outputFile = argv[3], configFile = argv[4];
    if (!isNoDummyEdge) {
        std::cerr << "Only -r 1 (no power-of-two dummy edges) is supported by the B200 engine." << std::endl;
        return -1;
    }
    const int T = (int)threadCount;
    const bool loopback = getenv("COGNN_B200_PLANE") && std::string(getenv("COGNN_B200_PLANE")) == "loopback";

    GNNConfig cfg;
    std::string err;
    if (!cfg.read(configFile, &err)) {
        std::cerr << err << std::endl;
        return -1;
    }

    // ---- partition file: vid tid (graph_io_util.h:58-87) ----
    std::vector<int64_t> tid;
    {
        std::ifstream in(partitionFile);
        if (!in.is_open()) {
            std::cerr << "Invalid format in graph topology input files." << std::endl;
            return -1;
        }
        std::string line;
        while (nextEffectiveLine(in, line)) {
            std::istringstream iss(line);
            uint64_t v;
            uint32_t t;
            if (!(iss >> v >> t)) {
                std::cerr << "Invalid format in graph topology input files." << std::endl;
                return -1;
            }
            t /= (uint32_t)(graphTileCount / threadCount);  // tile merge factor
            if (t >= (uint32_t)T) {
                std::cerr << "Invalid format in graph topology input files." << std::endl;
                return -1;
            }
            if (v >= tid.size()) tid.resize(v + 1, -1);
            tid[v] = t;
        }
        for (auto t : tid)
            if (t < 0) {
                std::cerr << "partition file must list every vertex id 0..N-1" << std::endl;
                return -1;
            }
    }
    const size_t N = tid.size();
    // ---- edge list: src dst [weight] (graph_io_util.h:121-165) ----
    std::vector<int64_t> edges;
    {
        std::ifstream in(edgelistFile);
        if (!in.is_open()) {
            std::cerr << "Invalid format in graph topology input files." << std::endl;
            return -1;
        }
        std::string line;
        while (nextEffectiveLine(in, line)) {
            char* pend = nullptr;
            const char* pbegin = line.c_str();
            uint64_t s = strtoull(pbegin, &pend, 10);
            if (pend == pbegin) {
                std::cerr << "Invalid format in graph topology input files." << std::endl;
                return -1;
            }
            pbegin = pend;
            uint64_t d = strtoull(pbegin, &pend, 10);
            if (pend == pbegin || s >= N || d >= N) {
                std::cerr << "Invalid format in graph topology input files." << std::endl;
                return -1;
            }
            edges.push_back((int64_t)s);
            edges.push_back((int64_t)d);
            if (undirected) {
                edges.push_back((int64_t)d);
                edges.push_back((int64_t)s);
            }
        }
    }
    // ---- vertex data: vid f_0 .. f_{F-1} label (harness.cpp:21-48, kernel_harness.h:37-44) ----
    const size_t F = cfg.input_dim;
    std::vector<double> feats(N * F, 0.0);
    std::vector<int32_t> labels(N, 0);
    {
        std::ifstream in(vertexlistFile);
        if (!in.is_open()) {
            std::cerr << "cannot open " << vertexlistFile << std::endl;
            return -1;
        }
        std::string line;
        while (nextEffectiveLine(in, line)) {
            std::istringstream iss(line);
            uint64_t v;
            if (!(iss >> v) || v >= N) {
                std::cerr << "bad vertex line in " << vertexlistFile << std::endl;
                return -1;
            }
            for (size_t j = 0; j < F; ++j) iss >> feats[v * F + j];
            iss >> labels[v];
        }
    }
    std::cout << "Graph loaded from " << edgelistFile << " and " << partitionFile << " with " << graphTileCount << " graph tiles, into "
              << threadCount << " tiles. Treated as " << (undirected ? "undirected" : "directed") << " graph.Current tile is the No."
              << tileIndex << " tile." << std::endl;

    // ---- engine ----
    int n_dev = cgb_device_count();
    if (n_dev <= 0) {
        std::cerr << "no CUDA device: the B200 engine has no CPU fallback" << std::endl;
        return -1;
    }
    const int device = getenv("COGNN_B200_DEVICE") ? atoi(getenv("COGNN_B200_DEVICE")) : (loopback ? 0 : (int)(tileIndex % (size_t)n_dev));
    cgb_ctx* ctx = nullptr;
    if (cgb_ctx_create(device, &ctx) != CGB_OK) {
        std::cerr << cgb_last_error(nullptr) << std::endl;
        return -1;
    }
    try {
        std::unique_ptr<Comm> comm;
        if (loopback) {
            comm = make_loopback_comm(T, ctx);
        } else {
            // rendezvous of the T independently launched processes: party 0 publishes the NCCL id next to the result file
            const std::string idfile = outputFile + ".ncclid";
            char uid[128];
            if (tileIndex == 0) {
                nccl_get_unique_id(uid);
                std::ofstream o(idfile + ".tmp", std::ios::binary);
                o.write(uid, 128);
                o.close();
                rename((idfile + ".tmp").c_str(), idfile.c_str());
            } else {
                for (int tries = 0;; ++tries) {
                    std::ifstream i(idfile, std::ios::binary);
                    if (i.is_open() && i.read(uid, 128)) break;
                    if (tries > 6000) throw std::runtime_error("timed out waiting for " + idfile);
                    std::this_thread::sleep_for(std::chrono::milliseconds(10));
                }
            }
            comm = make_nccl_comm((int)tileIndex, T, ctx, uid);
            if (tileIndex == 0) remove(idfile.c_str());
        }
        // Master key of the dealer emulation.  Every party derives ALL correlated randomness (OM masks, Beaver triples, re-share
        // masks) from it, so a run is only as private as this key is secret from ... nobody: it is a TEST configuration.  The key
        // comes from COGNN_B200_KEY (eight 32-bit words, hex, comma separated; the same for all parties of a run) -- there is no
        // compiled-in default any more -- and the binary refuses to start unless the caller acknowledges what it emulates.
        uint32_t key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        {
            const char* ack = getenv("COGNN_B200_ALLOW_INSECURE_EMULATION");
            const char* ks = getenv("COGNN_B200_KEY");
            printf("gcn-optimize-b200: INSECURE -- all correlated randomness comes from a trusted-dealer EMULATION keyed by COGNN_B200_KEY (every\n"
                   "party can regenerate every mask), and ReLU / softmax / ReLU' run as ideal-functionality stand-ins on values opened to the\n"
                   "owner.  Timing and functional parity only; no privacy.  (DESIGN.md section 1)\n");
            if (!ack || std::string(ack) != "1") {
                printf("refusing to run: set COGNN_B200_ALLOW_INSECURE_EMULATION=1 to acknowledge\n");
                exit(-1);
            }
            unsigned long long v[8] = {0};
            if (!ks || sscanf(ks, "%llx,%llx,%llx,%llx,%llx,%llx,%llx,%llx", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) != 8) {
                printf("refusing to run: COGNN_B200_KEY must hold eight comma-separated 32-bit hex words (e.g. from /dev/urandom)\n");
                exit(-1);
            }
            for (int i = 0; i < 8; ++i) key[i] = (uint32_t)v[i];
        }
        SSGcnEngine engine(comm.get(), cfg, CGB_SCALER_BITS, key);
        engine.verbose = true;
        std::cout << tileIndex << " Initialize graph algo kernel" << std::endl;
        auto t_pre = std::chrono::high_resolution_clock::now();
        for (int p = 0; p < T; ++p) {
            if (!comm->is_local(p)) continue;
            // graph tile + index vectors (graph_io_util.h:40-208, ssk.h:295-534) by device passes; COGNN_B200_INGEST=host keeps
            // the host builder
            const char* ing = getenv("COGNN_B200_INGEST");
            PartyGraph g = (ing && std::string(ing) == "host") ? build_party_graph(edges.data(), edges.size() / 2, tid.data(), N, T, p)
                                                               : build_party_graph_device(ctx, edges.data(), edges.size() / 2, tid.data(), N, T, p);
            std::vector<double> fl(g.vids.size() * F);
            std::vector<int32_t> ll(g.vids.size());
            for (size_t i = 0; i < g.vids.size(); ++i) {
                memcpy(&fl[i * F], &feats[g.vids[i] * F], F * sizeof(double));
                ll[i] = labels[g.vids[i]];
            }
            engine.add_party(g, fl.data(), ll.data());
        }
        printf("::preprocess took %lf seconds\n", std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t_pre).count());
        std::cout << tileIndex << " Begin vertex data sharing" << std::endl;
        engine.setup();
        std::cout << tileIndex << " Begin algo kernel iteration" << std::endl;
        engine.run(maxIters);
        std::ofstream res(outputFile);
        for (const auto& m : engine.metrics())
            res << "iter " << m.iter << " party " << m.party << " loss " << m.loss << " acc_full " << m.acc_full << " acc_train " << m.acc_train
                << " acc_test " << m.acc_test << "\n";
        std::cout << tileIndex << " Finish algo kernel" << std::endl;
    } catch (const std::exception& e) {
        printf("%s\n", e.what());  // the reference's error behaviour: print and exit(-1) (ssk.h:794-797)
        exit(-1);
    }
    cgb_ctx_destroy(ctx);
    return 0;
}
