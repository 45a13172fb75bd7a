// cli.cpp -- `minbpe-cc`: flag-compatible with the reference CLI (code/examples/minbpe-cc.cpp:96-131), backed by
// the B200 engine. CLI11 is not available in this image, so the flags are parsed by hand; long options accept
// both "--opt value" and "--opt=value". Extra flags: --engine stepwise|persistent, --device N, --threads N, --gpus N,
// --vocab-format reference|karpathy.
//
// --gpus N (N > 1) runs ONE PROCESS PER GPU, as the library's multi-GPU paths want it: the program starts N-1 copies of
// itself (ranks 1..N-1, devices device+1..), all ranks cut the input at the same regex-safe points (after a newline,
// before a printable ASCII byte: SURVEY H7) and take one part each.
//   --train   mbpe_train_text_sharded: split + dedup per rank, one all-gather of the unique chunks over NCCL, merge loop;
//             rank 0 writes the model
//   --encode  no communication: every rank encodes its part into <output>.partR, rank 0 concatenates them in order
//   --decode  the id stream in N equal ranges, same hand-over
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "bpe_host.hpp"

using namespace mbpe::host;

static bool slurp(const std::string &path, std::string &out, std::string &err) { // minbpe-cc.cpp:36-46
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        err = std::strerror(errno);
        return false;
    }
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    f.seekg(0);
    out.resize(n > 0 ? (size_t)n : 0);
    if (n > 0) f.read(out.data(), n);
    return true;
}

static void usage() {
    std::cout << "Training, encoding and decoding of tokens\nUsage: minbpe-cc [OPTIONS]\n\nOptions:\n"
                 "  -h,--help                   Print this help message and exit\n"
                 "  -i,--input TEXT             Path to the input to be trained on, encoded or decoded\n"
                 "  -o,--output TEXT            Path for the output of the encoding or decoding\n"
                 "  -s,--special-tokens-path TEXT\n                              Path to the special tokens file\n"
                 "  -t,--train                  Train on the input\n"
                 "  -d,--decode                 Decode the input\n"
                 "  -e,--encode                 Encode the input\n"
                 "  -w,--write-vocab            When training, write the vocabulary to a file\n"
                 "  --vocab-size INT            Vocabulary size\n"
                 "  --encoder TEXT              Encoder to use from basic,gpt2,gpt4\n"
                 "  -m,--model-path TEXT        Path to load or save the model\n"
                 "  -v,--verbose                Print more things\n"
                 "  -c,--conflict-resolution TEXT:{first,lexical}\n"
                 "                              Conflict resolution strategy: 'first' or 'lexical'\n"
                 "  --engine TEXT:{persistent,stepwise}  how the GPU merge loop is driven\n"
                 "  --device INT                CUDA device (first device with --gpus)\n"
                 "  --gpus INT                  number of GPUs, one process per GPU (train: sharded front end over NCCL;\n"
                 "                              encode / decode: the input in contiguous parts, no communication)\n"
                 "  --vocab-format TEXT:{reference,karpathy}  layout of the .vocab file written with -w\n"
                 "  --threads INT               host pre-tokenisation threads (0 = all)\n";
}

static double cli_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    const double t_main = cli_now();
    std::string input_path, output_path, special_path, encoder = "gpt4", model_path = "./output.model";
    std::string conflict = "first", engine = "persistent", vocab_format = "reference", rendezvous;
    bool train = false, decode = false, encode = false, write_vocab = false, verbose = false;
    int vocab_size = 512, device = 0, threads = 0, gpus = 1, rank = 0;

    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], val;
        bool has_val = false;
        if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            if (eq != std::string::npos) {
                val = a.substr(eq + 1);
                a = a.substr(0, eq);
                has_val = true;
            }
        }
        auto value = [&]() -> std::string {
            if (has_val) return val;
            if (i + 1 >= argc) {
                std::cerr << a << ": 1 required TEXT missing\nRun with --help for more information.\n";
                exit(106);
            }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") {
            usage();
            return 0;
        } else if (a == "-i" || a == "--input") input_path = value();
        else if (a == "-o" || a == "--output") output_path = value();
        else if (a == "-s" || a == "--special-tokens-path") special_path = value();
        else if (a == "-t" || a == "--train") train = true;
        else if (a == "-d" || a == "--decode") decode = true;
        else if (a == "-e" || a == "--encode") encode = true;
        else if (a == "-w" || a == "--write-vocab") write_vocab = true;
        else if (a == "-v" || a == "--verbose") verbose = true;
        else if (a == "--vocab-size") vocab_size = std::atoi(value().c_str());
        else if (a == "--encoder") encoder = value();
        else if (a == "-m" || a == "--model-path") model_path = value();
        else if (a == "-c" || a == "--conflict-resolution") {
            conflict = value();
            if (conflict != "first" && conflict != "lexical") {
                std::cerr << "--conflict-resolution: " << conflict << " not in {first,lexical}\n"
                          << "Run with --help for more information.\n";
                return 105;
            }
        } else if (a == "--engine") engine = value();
        else if (a == "--device") device = std::atoi(value().c_str());
        else if (a == "--threads") threads = std::atoi(value().c_str());
        else if (a == "--gpus") gpus = std::max(1, std::atoi(value().c_str()));
        else if (a == "--vocab-format") {
            vocab_format = value();
            if (vocab_format != "reference" && vocab_format != "karpathy") {
                std::cerr << "--vocab-format: " << vocab_format << " not in {reference,karpathy}\nRun with --help for more information.\n";
                return 105;
            }
        } else if (a == "--mbpe-rank") rank = std::atoi(value().c_str());          // (set by the launcher below, not by users)
        else if (a == "--mbpe-rendezvous") rendezvous = value();
        else {
            std::cerr << "The following argument was not expected: " << a << "\nRun with --help for more information.\n";
            return 109;
        }
    }

    if (input_path.empty()) { // minbpe-cc.cpp:135-144
        std::cerr << "Input file not specified\n";
        return -1;
    }
    if (!std::filesystem::exists(input_path)) {
        std::cerr << "Input file " << input_path << " does not exist\n";
        return -1;
    }
    // ---- --gpus N: one process per GPU. Nothing has touched CUDA yet, so starting copies of this program is safe. ----
    std::vector<pid_t> children;
    if (gpus > 1 && rank == 0 && rendezvous.empty()) {
        if (encoder == "basic" && !decode) {
            std::cerr << "--gpus needs a regex encoder (gpt2 / gpt4): with `basic` the whole text is one chunk\n";
            return -1;
        }
        char tmpl[] = "/tmp/minbpe-cc-XXXXXX";
        if (!mkdtemp(tmpl)) {
            std::cerr << "cannot create a rendezvous directory: " << std::strerror(errno) << "\n";
            return 1;
        }
        rendezvous = tmpl;
        for (int r = 1; r < gpus; r++) {
            pid_t pid = fork();
            if (pid == 0) {
                std::vector<std::string> args(argv, argv + argc);
                args.push_back("--mbpe-rank=" + std::to_string(r));
                args.push_back("--mbpe-rendezvous=" + rendezvous);
                std::vector<char *> av;
                for (auto &x : args) av.push_back(x.data());
                av.push_back(nullptr);
                execv("/proc/self/exe", av.data());
                _exit(127);
            }
            if (pid < 0) {
                std::cerr << "fork failed: " << std::strerror(errno) << "\n";
                return 1;
            }
            children.push_back(pid);
        }
    }
    device += rank;
    auto wait_children = [&]() { // rank 0: collect the other ranks, remove the rendezvous directory
        int bad = 0;
        for (pid_t pid : children) {
            int st = 0;
            if (waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
        }
        if (rank == 0 && !rendezvous.empty()) {
            std::error_code ec;
            std::filesystem::remove_all(rendezvous, ec);
        }
        return bad;
    };
    // hand-over through the rendezvous directory: a file appears when its content is complete (rename is atomic)
    auto publish = [&](const std::string &name, const std::string &content) {
        std::ofstream(rendezvous + "/" + name + ".tmp", std::ios::binary) << content;
        std::filesystem::rename(rendezvous + "/" + name + ".tmp", rendezvous + "/" + name);
    };
    auto await = [&](const std::string &name, std::string &content) -> bool {
        for (int i = 0; i < 60000; i++) { // 10 minutes
            std::string e2;
            if (std::filesystem::exists(rendezvous + "/" + name)) return slurp(rendezvous + "/" + name, content, e2);
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        return false;
    };
    // byte range of this rank: the file in `gpus` parts cut at the first regex-safe point at or after r * size / gpus
    // (every rank computes the same cuts); unit = 4 for the id stream of --decode (no boundary rule there)
    auto my_range = [&](uint64_t &begin, uint64_t &end) -> bool {
        std::error_code ec;
        const uint64_t size = std::filesystem::file_size(input_path, ec);
        if (ec) return false;
        std::vector<uint64_t> cut(gpus + 1, size);
        cut[0] = 0;
        std::ifstream f(input_path, std::ios::binary);
        for (int k = 1; k < gpus; k++) {
            uint64_t t = size / gpus * k;
            if (decode) {
                cut[k] = t / 4 * 4;
                continue;
            }
            t = std::max(t, cut[k - 1]);
            uint64_t found = size;
            std::vector<char> win(1 << 20);
            for (uint64_t at = t ? t - 1 : 0; at < size && found == size; at += win.size() - 1) {
                f.clear();
                f.seekg((std::streamoff)at);
                f.read(win.data(), (std::streamsize)win.size());
                const size_t got = (size_t)f.gcount();
                for (size_t i = 1; i < got; i++)
                    if (win[i - 1] == '\n' && (unsigned char)win[i] >= 0x21 && (unsigned char)win[i] <= 0x7E && at + i >= t) {
                        found = at + i;
                        break;
                    }
                if (got < win.size()) break;
            }
            cut[k] = found;
        }
        begin = cut[rank];
        end = cut[rank + 1];
        return true;
    };
    auto slurp_range = [&](uint64_t begin, uint64_t end, std::string &out) -> bool {
        std::ifstream f(input_path, std::ios::binary);
        if (!f) return false;
        out.resize(end - begin);
        f.seekg((std::streamoff)begin);
        if (end > begin) f.read(out.data(), (std::streamsize)(end - begin));
        return (uint64_t)f.gcount() == end - begin;
    };
    std::string special_data, err;
    bool have_special = false;
    if (!special_path.empty() && std::filesystem::exists(special_path)) { // minbpe-cc.cpp:149-159
        if (slurp(special_path, special_data, err)) {
            have_special = true;
            std::cout << "Loaded special tokens from " << special_path << "\n";
        } else {
            std::cerr << "Failed to load special tokens from " << special_path << ": " << err << "\n";
        }
    }
    auto t1 = std::chrono::high_resolution_clock::now();

    std::string pattern;
    if (encoder == "gpt2") pattern = kGpt2Pattern;
    else if (encoder == "gpt4") pattern = kGpt4Pattern;
    else if (encoder == "basic") pattern = "";
    else {
        std::cout << "Encoder should be one of: basic, gpt2 or gpt4\n";
        return -1;
    }
    if (mbpe_device_count() == 0) {
        std::cerr << "No CUDA device found: this build of minbpe-cc runs its merge loops on a B200 and has no CPU path\n";
        return 3;
    }
    Tokenizer rt(pattern, device);
    rt.set_engine(engine == "stepwise" ? MBPE_ENGINE_STEPWISE : MBPE_ENGINE_PERSISTENT);
    rt.set_threads(threads);
    int rc = 0;

    if (gpus > 1) { // one process per GPU (see the head of this file); this process is rank `rank` on device `device`
        uint64_t begin = 0, end = 0;
        std::string part;
        if (!my_range(begin, end) || !slurp_range(begin, end, part)) {
            std::cerr << "Failed to read the input file\n";
            rc = MBPE_E_IO;
        }
        uint64_t n_out = 0;
        if (rc == 0 && train) {
            uint8_t id[128];
            std::string ids;
            if (rank == 0) {
                if ((rc = mbpe_comm_unique_id(id)) == 0) publish("nccl_id", std::string(reinterpret_cast<char *>(id), 128));
            } else if (await("nccl_id", ids) && ids.size() == 128) {
                memcpy(id, ids.data(), 128);
            } else {
                rc = MBPE_E_IO;
            }
            mbpe_comm *comm = nullptr;
            mbpe_pretok *pt = nullptr;
            if (rc == 0) rc = mbpe_comm_create(id, rank, gpus, device, &comm);
            if (rc == 0) rc = mbpe_pretok_create(device, &pt);
            if (rc == 0) rc = mbpe_pretok_select(pt, pattern.c_str());
            const uint32_t n_target = (uint32_t)std::max(vocab_size - 256, 0);
            std::vector<uint32_t> m(2ull * std::max<uint32_t>(n_target, 1));
            std::vector<int32_t> counts(std::max<uint32_t>(n_target, 1));
            uint32_t n_merges = 0;
            if (rank == 0)
                std::cout << "Training using file \"" << input_path << "\" encoder " << encoder << " vocab size " << vocab_size
                          << " model path " << model_path << " on " << gpus << " GPUs\n";
            if (rc == 0)
                rc = mbpe_train_text_sharded(comm, pt, reinterpret_cast<const uint8_t *>(part.data()), part.size(), (uint32_t)vocab_size,
                                             conflict == "first" ? MBPE_MODE_FIRST : MBPE_MODE_LEXICAL, m.data(), counts.data(),
                                             &n_merges, nullptr, nullptr);
            if (rc == 0 && rank == 0) { // every rank holds the same merge list; rank 0 writes it (Tokenizer.h:875-926)
                rc = mbpe_write_model(model_path.c_str(), pattern.c_str(), have_special ? special_data.data() : nullptr,
                                      have_special ? special_data.size() : 0, m.data(), n_merges, write_vocab && vocab_format == "reference");
                if (rc == 0 && write_vocab && vocab_format == "karpathy")
                    rc = mbpe_write_vocab_karpathy((model_path + ".vocab").c_str(), have_special ? special_data.data() : nullptr,
                                                   have_special ? special_data.size() : 0, m.data(), n_merges);
            }
            if (rc) std::cerr << "Training failed (rank " << rank << "): " << mbpe_last_error() << "\n";
            if (pt) mbpe_pretok_destroy(pt);
            if (comm) mbpe_comm_destroy(comm);
        } else if (rc == 0 && (encode || decode)) {
            if (output_path.empty()) {
                std::cerr << "Output file not specified\n";
                rc = MBPE_E_INVALID;
            }
            if (rc == 0 && have_special) rt.set_special_tokens_from_file(special_data); // (load() adds the model's own)
            if (rc == 0) rc = rt.load(model_path, verbose && rank == 0);
            const std::string part_path = rank == 0 ? output_path : output_path + ".part" + std::to_string(rank);
            if (rc == 0 && encode) {
                std::vector<Token> ids;
                rc = rt.encode(part, false, ids);
                if (rc == 0) {
                    std::ofstream f(part_path, std::ios::binary);
                    f.write(reinterpret_cast<const char *>(ids.data()), (std::streamsize)(ids.size() * sizeof(Token)));
                    n_out = ids.size();
                    if (!f) rc = MBPE_E_IO;
                }
            } else if (rc == 0) {
                std::vector<Token> ids(part.size() / sizeof(Token));
                std::memcpy(ids.data(), part.data(), ids.size() * sizeof(Token));
                std::string text;
                rc = rt.decode(ids, false, text);
                if (rc == 0) {
                    std::ofstream f(part_path, std::ios::binary);
                    f << text;
                    n_out = text.size();
                    if (!f) rc = MBPE_E_IO;
                }
            }
            if (rc) std::cerr << (encode ? "Encoding" : "Decoding") << " failed (rank " << rank << "): " << rt.error() << "\n";
            if (rank != 0) {
                publish("done." + std::to_string(rank), rc ? "fail" : std::to_string(n_out));
            } else { // append the other ranks' parts in order
                std::ofstream out(output_path, std::ios::binary | std::ios::app);
                for (int r = 1; r < gpus && rc == 0; r++) {
                    std::string st;
                    if (!await("done." + std::to_string(r), st) || st == "fail") {
                        rc = MBPE_E_IO;
                        break;
                    }
                    n_out += std::strtoull(st.c_str(), nullptr, 10);
                    const std::string pp = output_path + ".part" + std::to_string(r);
                    std::ifstream in(pp, std::ios::binary);
                    out << in.rdbuf();
                    in.close();
                    std::error_code ec;
                    std::filesystem::remove(pp, ec);
                }
                if (rc == 0) std::cout << "Writing " << n_out << (encode ? " encoded tokens\nSuccess\n" : " decoded bytes\n");
            }
        }
        if (rank == 0) {
            rc |= wait_children();
            auto t2 = std::chrono::high_resolution_clock::now();
            std::cout << "Execution time: " << std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count() / 1000.0
                      << " (s)" << std::endl;
        }
        return rc ? 1 : 0;
    }

    if (train) { // minbpe-cc.cpp:181-211
        if (have_special) rt.set_special_tokens_from_file(special_data);
        else std::cout << "No special tokens file provided\n";
        std::cout << "Training using file \"" << input_path << "\" encoder " << encoder << " vocab size " << vocab_size
                  << " model path " << model_path << "\n";
        if (verbose) std::cout << "Loading file " << input_path << "\n";
        std::string text;
        if (slurp(input_path, text, err)) {
            if (verbose) std::cout << "Starting training...\n";
            rc = rt.train(text, vocab_size, conflict == "first" ? Tokenizer::FIRST : Tokenizer::LEXICAL, verbose);
            if (rc == 0) rc = rt.save(model_path, write_vocab && vocab_format == "reference");
            if (rc == 0 && write_vocab && vocab_format == "karpathy") rc = rt.save_vocab_karpathy(model_path + ".vocab");
            if (rc) std::cerr << "Training failed: " << rt.error() << "\n";
        } else {
            std::cerr << "Failed to load training input file: " << err << "\n";
        }
    } else if (encode) { // minbpe-cc.cpp:212-242
        if (output_path.empty()) {
            std::cerr << "Output file not specified\n";
            return -1;
        }
        if (!std::filesystem::exists(model_path)) {
            std::cerr << "Model file " << model_path << " does not exist\n";
            return -1;
        }
        std::cout << "Encoding input file \"" << input_path << "\" encoder " << encoder << " model path " << model_path
                  << " output to " << output_path << "\n";
        rc = rt.load(model_path, verbose);
        if (getenv("MBPE_DEBUG")) fprintf(stderr, "[mbpe] cli: model loaded %.0f ms after start\n", (cli_now() - t_main) * 1e3);
        std::string text;
        uint64_t n_streamed = 0;
        std::error_code size_ec;
        const bool big = rc == 0 && special_path.empty() && !getenv("MBPE_CLI_NO_STREAM") && std::filesystem::file_size(input_path, size_ec) >= (8u << 20) && !size_ec;
        if (big && (rc = rt.encode_file(input_path, output_path, &n_streamed)) == 0) {
            // block-wise: disk -> pinned memory -> GPU -> pinned memory -> disk, never the whole file in memory
            std::cout << "Writing " << n_streamed << " encoded tokens\nSuccess\n";
        } else if (big && rc != MBPE_E_UNSUPPORTED) {
            std::cerr << "Encoding failed: " << rt.error() << "\n";
        } else if ((rc = (rc == MBPE_E_UNSUPPORTED ? 0 : rc)) == 0 && slurp(input_path, text, err)) {
            std::vector<Token> ids;
            rc = rt.encode(text, verbose, ids);
            if (rc == 0) {
                std::cout << "Writing " << ids.size() << " encoded tokens\n";
                std::ofstream f(output_path, std::ios::binary); // raw little-endian u32 stream (minbpe-cc.cpp:58-69)
                if (f) {
                    f.write(reinterpret_cast<const char *>(ids.data()), (std::streamsize)(ids.size() * sizeof(Token)));
                    std::cout << "Success\n";
                } else {
                    std::cerr << "Failed with error: " << std::strerror(errno) << "\n";
                }
            } else {
                std::cerr << "Encoding failed: " << rt.error() << "\n";
            }
        } else if (rc == 0) {
            std::cerr << "Failed with error: " << err << "\n";
        }
    } else if (decode) { // minbpe-cc.cpp:243-258
        std::cout << "Decoding input file \"" << input_path << "\" encoder " << encoder << " model path " << model_path
                  << " output to " << output_path << "\n";
        rc = rt.load(model_path, verbose);
        std::string raw;
        std::error_code dsize_ec;
        const bool big_enc = rc == 0 && !verbose && !getenv("MBPE_CLI_NO_STREAM") &&
                             std::filesystem::file_size(input_path, dsize_ec) >= (8u << 20) && !dsize_ec;
        uint64_t s_ids = 0, s_bytes = 0;
        if (big_enc) { // block-wise: disk -> pinned memory -> GPU -> pinned memory -> disk
            rc = rt.decode_file(input_path, output_path, &s_ids, &s_bytes);
            if (rc == 0)
                std::cout << "Loaded encoding with " << s_ids << " tokens\nWriting " << s_bytes << " decoded tokens to "
                          << output_path << "\n";
            else
                std::cerr << "Decoding failed: " << rt.error() << "\n";
        } else if (rc == 0 && slurp(input_path, raw, err)) {
            std::vector<Token> ids(raw.size() / sizeof(Token)); // trailing partial word is dropped (minbpe-cc.cpp:79)
            std::memcpy(ids.data(), raw.data(), ids.size() * sizeof(Token));
            std::cout << "Loaded encoding with " << ids.size() << " tokens\n";
            std::string text;
            rc = rt.decode(ids, verbose, text);
            if (rc == 0) {
                std::cout << "Writing " << text.size() << " decoded tokens to " << output_path << "\n";
                std::ofstream f(output_path, std::ios::binary);
                f << text;
            } else {
                std::cerr << "Decoding failed: " << rt.error() << "\n";
            }
        } else if (rc == 0) {
            std::cerr << "Failed with error: " << err << "\n";
        }
    }
    auto t2 = std::chrono::high_resolution_clock::now();
    std::cout << "Execution time: " << std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count() / 1000.0
              << " (s)" << std::endl;
    return rc ? 1 : 0;
}
