// compress.cpp — host driver of the B200 shared_tree path.
//
// Same command line as the reference's compress.cpp (flags at compress.cpp:82-93 of the
// reference: --help --verbose --statistics --no-save --output= --histogram= --dna-size=) and
// the same --statistics CSV line, extended with device throughput; plus --decompress, which
// the reference lacks (its .dag format does not record dna::size, so it must be given).
// Everything below the flag parsing is calls into the shim in include/ (one C-ABI call each).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>

#include "shared_tree.h"

namespace fs = std::filesystem;
using clk = std::chrono::steady_clock;

struct Options {
  fs::path input, output, histogram;
  bool verbose = false, statistics = false, save = true, decompress = false;
  std::size_t dna_size = 12;
};

static void usage() {
  std::cout << "Usage: compress_b200 [options] file\n"
               "  --help               this text\n"
               "  --verbose            human-readable report\n"
               "  --statistics         one CSV line: dna_size,width,ratio,in_bytes,out_bytes,construct_ms,sort_ms,total_ms,build_gbp_s\n"
               "  --no-save            do not write the .dag file\n"
               "  --output=<file>      output path (default <input>.dag, or <input>.txt with --decompress)\n"
               "  --histogram=<file>   write per-layer reference-count histograms (CSV)\n"
               "  --dna-size=<n>       nucleotides per leaf, 1..16 (default 12)\n"
               "  --decompress         input is a .dag file; write the decoded sequence as text\n";
}

static Options parse(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto value = [&](const char* key) { return a.substr(std::strlen(key)); };
    if (a == "--help") { usage(); std::exit(0); }
    else if (a == "--verbose") o.verbose = true;
    else if (a == "--statistics") o.statistics = true;
    else if (a == "--no-save") o.save = false;
    else if (a == "--decompress") o.decompress = true;
    else if (a.rfind("--output=", 0) == 0) o.output = value("--output=");
    else if (a.rfind("--histogram=", 0) == 0) o.histogram = value("--histogram=");
    else if (a.rfind("--dna-size=", 0) == 0) o.dna_size = (std::size_t)std::atoi(value("--dna-size=").c_str());
    else if (o.input.empty()) o.input = a;
    else { std::cout << "Compression of multiple files at once is currently not supported.\n"; std::exit(1); }
  }
  if (o.verbose && o.statistics) { std::cout << "Invalid flag combination: --verbose and --statistics are mutually exclusive\n"; std::exit(2); }
  if (o.input.empty()) { std::cout << "Invalid command: argument <file> required.\nUse --help for more information\n"; std::exit(2); }
  if (o.dna_size < 1 || o.dna_size > 16) { std::cout << "Invalid --dna-size: 1..16 supported\n"; std::exit(2); }
  if (!fs::is_regular_file(o.input)) { std::cout << "Invalid filename: " << o.input << '\n'; std::exit(2); }
  if (o.output.empty() && o.save) { o.output = o.input; o.output.replace_extension(o.decompress ? ".txt" : ".dag"); }
  return o;
}

static double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

static int decompress(const Options& o) {
  std::ifstream in{o.input, std::ios::binary};
  shared_tree tree = shared_tree::deserialize(in);
  const std::size_t width = tree.width();
  std::string text(width * dna::size(), '\0');
  const int st = stb_decode_ascii(tree.handle(), 0, width, text.data(), STB_HOST);
  if (st != STB_OK) { std::cerr << stb_last_error(tree.handle()) << '\n'; return 1; }
  if (o.save) std::ofstream{o.output, std::ios::binary}.write(text.data(), (std::streamsize)text.size());
  if (o.verbose || o.statistics) std::cout << dna::size() << ',' << width << ',' << text.size() << '\n';
  return 0;
}

int main(int argc, char** argv) {
  const Options o = parse(argc, argv);
  dna::size(o.dna_size);
  if (o.decompress) return decompress(o);
  const auto in_bytes = fs::file_size(o.input);
  if (o.verbose) std::cout << " Filename: " << o.input << "\n Size:     " << bytes_to_string(in_bytes) << "\n";

  auto t0 = clk::now();
  shared_tree tree{o.input, o.verbose};
  const double construct_ms = ms_since(t0);
  t0 = clk::now();
  tree.sort_tree(o.verbose);
  const double sort_ms = ms_since(t0);

  const std::size_t out_bytes = tree.bytes(), width = tree.width();
  if (!o.histogram.empty()) tree.store_histogram(o.histogram);
  if (!o.output.empty() && o.save) tree.save(o.output);

  const double gbps = double(width * dna::size()) / (construct_ms * 1e6);
  if (o.verbose) {
    std::cout << " Output:            " << o.output << "\n Size:              " << bytes_to_string(out_bytes)
              << "\n Nucleotides:       " << width * dna::size() << "\n Compression ratio: " << double(in_bytes) / double(out_bytes)
              << "\n Leaf size:         " << dna::size() << "\n Width:             " << width << "\n Depth:             " << tree.depth()
              << "\n Leaves:            " << tree.leaf_count() << "\n Nodes:             " << tree.node_count()
              << "\n Tree construction: " << construct_ms << " ms (" << gbps << " Gbp/s incl. file read + H2D)\n Frequency sorting: " << sort_ms << " ms\n";
  }
  if (o.statistics)
    std::cout << dna::size() << ',' << width << ',' << double(in_bytes) / double(out_bytes) << ',' << in_bytes << ',' << out_bytes << ','
              << (long long)construct_ms << ',' << (long long)sort_ms << ',' << (long long)(construct_ms + sort_ms) << ',' << gbps << '\n';
  return 0;
}
