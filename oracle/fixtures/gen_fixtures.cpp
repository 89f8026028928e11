// oracle/fixtures/gen_fixtures.cpp — TEST INFRASTRUCTURE ONLY.
// Materialises the reference's tropical fixtures fst_000..fst_020 (rustfst-tests-data/fst_NNN/fst_NNN.h,
// included in place from /root/reference via -I, never copied) as OpenFst binary vector files:
//   <out>/fst_NNN_raw.fst      = get_fst()          with all property bits computed (main.cpp:1045-1048)
//   <out>/fst_NNN_compose.fst  = get_fst_compose()  with all property bits computed (main.cpp:1191-1193)
// fst_005, fst_010 (LogArc) and fst_011 (Tropical x Log) are outside the TropicalWeight scope.
// Run with cwd = /root/reference/rustfst-tests-data (the .fst.in paths are relative to it).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include "fst_shim.h"
#include "utils.h"
#include "fst_000/fst_000.h"
#include "fst_001/fst_001.h"
#include "fst_002/fst_002.h"
#include "fst_003/fst_003.h"
#include "fst_004/fst_004.h"
#include "fst_006/fst_006.h"
#include "fst_007/fst_007.h"
#include "fst_008/fst_008.h"
#include "fst_009/fst_009.h"
#include "fst_012/fst_012.h"
#include "fst_013/fst_013.h"
#include "fst_014/fst_014.h"
#include "fst_015/fst_015.h"
#include "fst_016/fst_016.h"
#include "fst_017/fst_017.h"
#include "fst_018/fst_018.h"
#include "fst_019/fst_019.h"
#include "fst_020/fst_020.h"

template <class F>
void dump(const F& data, const std::string& name, const std::string& out) {
  auto raw = data.get_fst();
  raw.impl.props = oracle::compute_fst_properties_all(raw.impl);
  oracle::write_file(out + "/" + name + "_raw.fst", oracle::fst_to_bytes(raw.impl));
  auto comp = data.get_fst_compose();
  comp.impl.props = oracle::compute_fst_properties_all(comp.impl);
  oracle::write_file(out + "/" + name + "_compose.fst", oracle::fst_to_bytes(comp.impl));
  std::printf("%s raw: %zu states %zu arcs props=%016llx | compose: %zu states %zu arcs props=%016llx\n", name.c_str(),
              raw.impl.num_states(), raw.impl.num_trs_total(), (unsigned long long)raw.impl.props,
              comp.impl.num_states(), comp.impl.num_trs_total(), (unsigned long long)comp.impl.props);
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: gen_fixtures <out_dir>\n"); return 2; }
  std::string out = argv[1];
  dump(FstTestData000(), "fst_000", out);
  dump(FstTestData001(), "fst_001", out);
  dump(FstTestData002(), "fst_002", out);
  dump(FstTestData003(), "fst_003", out);
  dump(FstTestData004(), "fst_004", out);
  dump(FstTestData006(), "fst_006", out);
  dump(FstTestData007(), "fst_007", out);
  dump(FstTestData008(), "fst_008", out);
  dump(FstTestData009(), "fst_009", out);
  dump(FstTestData012(), "fst_012", out);
  dump(FstTestData013(), "fst_013", out);
  dump(FstTestData014(), "fst_014", out);
  dump(FstTestData015(), "fst_015", out);
  dump(FstTestData016(), "fst_016", out);
  dump(FstTestData017(), "fst_017", out);
  dump(FstTestData018(), "fst_018", out);
  dump(FstTestData019(), "fst_019", out);
  dump(FstTestData020(), "fst_020", out);
  return 0;
}
