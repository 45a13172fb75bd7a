// cli.cpp -- `minbpe-cc`: flag-compatible with the reference CLI (code/examples/minbpe-cc.cpp:96-131), backed by
// the B200 engine. CLI11 is not available in this image, so the flags are parsed by hand; long options accept
// both "--opt value" and "--opt=value". Extra flags: --engine stepwise|persistent, --device N, --threads N.
#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>

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
                 "  --device INT                CUDA device\n"
                 "  --threads INT               host pre-tokenisation threads (0 = all)\n";
}

static double cli_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    const double t_main = cli_now();
    std::string input_path, output_path, special_path, encoder = "gpt4", model_path = "./output.model";
    std::string conflict = "first", engine = "persistent";
    bool train = false, decode = false, encode = false, write_vocab = false, verbose = false;
    int vocab_size = 512, device = 0, threads = 0;

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
            if (rc == 0) rc = rt.save(model_path, write_vocab);
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
