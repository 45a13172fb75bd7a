// synthgen.cpp -- stand-alone writer of the benchmarks' synthetic corpus (TEST / BENCH INFRASTRUCTURE).
// bench.py's reference arm gets its input from this binary so that no product library is loaded in that process.
//   synthgen <seed> <n_bytes> <out_path> [first_block]
// Same bytes as mbpe_synth_corpus(seed, ...) of the library: both include minbpe-cc_b200/host/synth.hpp.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../minbpe-cc_b200/host/synth.hpp"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: synthgen <seed> <n_bytes> <out_path> [first_block]\n");
        return 2;
    }
    const uint64_t seed = strtoull(argv[1], nullptr, 0), n = strtoull(argv[2], nullptr, 0);
    const uint64_t first = argc > 4 ? strtoull(argv[4], nullptr, 0) : 0;
    std::vector<uint8_t> buf(n);
    mbpe::host::synth::generate(seed, first, buf.data(), n, 0);
    FILE *f = fopen(argv[3], "wb");
    if (!f || fwrite(buf.data(), 1, n, f) != n || fclose(f) != 0) {
        fprintf(stderr, "synthgen: cannot write %s\n", argv[3]);
        return 1;
    }
    return 0;
}
