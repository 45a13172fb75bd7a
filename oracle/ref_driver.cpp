// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin command-line driver around the UNMODIFIED reference headers
// (/root/reference/code/include/{Tokenizer,PairCount}.h), compiled against oracle/shim.
// It calls the same public methods, in the same order, as the reference CLI does
// (code/examples/minbpe-cc.cpp:179-207 train, :227-233 encode, :248-254 decode); CLI11 is
// not installed, so flags are positional here. Output goes to oracle/_ref/ only.
//
//   ref_driver train  <input> <model> <vocab> <basic|gpt2|gpt4> <first|lexical> [-s special] [-w] [-v]
//   ref_driver encode <input> <model> <out.enc>
//   ref_driver decode <input.enc> <model> <out.txt>
//
// "REF_TIME_S <seconds>" on stderr is the wall time of Tokenizer::train / ::encode / ::decode
// alone (no file I/O), used as the CPU baseline.
#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>  // Tokenizer.h uses std::setw without including it
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "Tokenizer.h"

using MinBpeCC::Tokenizer::Token;
using MinBpeCC::Tokenizer::Tokenizer;

static bool slurp(const std::string &p, std::string &out) {
    std::ifstream f(p, std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

static double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    if (argc < 5) {
        std::cerr << "usage: ref_driver train|encode|decode ...\n";
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd == "train") {
        if (argc < 7) return 2;
        std::string input = argv[2], model = argv[3];
        int vocab = std::atoi(argv[4]);
        std::string enc = argv[5], mode = argv[6], special;
        bool write_vocab = false, verbose = false;
        for (int i = 7; i < argc; i++) {
            if (!std::strcmp(argv[i], "-s") && i + 1 < argc)
                special = argv[++i];
            else if (!std::strcmp(argv[i], "-w"))
                write_vocab = true;
            else if (!std::strcmp(argv[i], "-v"))
                verbose = true;
        }
        std::string pattern;
        if (enc == "gpt2")
            pattern = Tokenizer::GPT2_SPLIT_PATTERN;
        else if (enc == "gpt4")
            pattern = Tokenizer::GPT4_SPLIT_PATTERN;
        else if (enc != "basic")
            return 2;
        Tokenizer rt(pattern);
        if (!special.empty()) {
            std::string sp;
            if (!slurp(special, sp)) return 3;
            rt.set_special_tokens_from_file(sp);
        }
        std::string text;
        if (!slurp(input, text)) return 3;
        auto cr = mode == "first" ? Tokenizer::CONFLICT_RESOLUTION::FIRST : Tokenizer::CONFLICT_RESOLUTION::LEXICAL;
        double t0 = now_s();
        rt.train(text, vocab, cr, verbose);
        double t1 = now_s();
        std::cerr << "REF_TIME_S " << (t1 - t0) << "\n";
        return rt.save(model, write_vocab) ? 0 : 4;
    }
    if (cmd == "encode") {
        std::string input = argv[2], model = argv[3], out = argv[4];
        Tokenizer rt(Tokenizer::GPT4_SPLIT_PATTERN);  // replaced by load(), as in the reference CLI
        if (!rt.load(model, false)) return 4;
        std::string text;
        if (!slurp(input, text)) return 3;
        double t0 = now_s();
        auto ids = rt.encode(text, false);
        double t1 = now_s();
        std::cerr << "REF_TIME_S " << (t1 - t0) << "\n";
        std::ofstream f(out, std::ios::binary);
        f.write(reinterpret_cast<const char *>(ids.data()), ids.size() * sizeof(Token));
        return f ? 0 : 4;
    }
    if (cmd == "decode") {
        std::string input = argv[2], model = argv[3], out = argv[4];
        Tokenizer rt(Tokenizer::GPT4_SPLIT_PATTERN);
        if (!rt.load(model, false)) return 4;
        std::string raw;
        if (!slurp(input, raw)) return 3;
        std::vector<Token> ids(raw.size() / sizeof(Token));
        std::memcpy(ids.data(), raw.data(), ids.size() * sizeof(Token));
        double t0 = now_s();
        auto text = rt.decode(ids, false);
        double t1 = now_s();
        std::cerr << "REF_TIME_S " << (t1 - t0) << "\n";
        std::ofstream f(out, std::ios::binary);
        f << text;
        return f ? 0 : 4;
    }
    return 2;
}
