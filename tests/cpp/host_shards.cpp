// The C++ driver's shard planner (ngs_b200/host/bam.hpp: read_bai, plan_shards, shard_bytes) on CPU.
// usage: host_shards <file.bam> <first_record_voffset> <n_refs> <n_shards>
// prints one line per shard: first_voffset end_voffset empty lo hi contigs...
#include <cstdio>
#include <cstdlib>

#include "../../ngs_b200/host/bam.hpp"

using namespace ngs;

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  try {
    MappedFile file(argv[1]);
    BaiIndex bai = read_bai(std::string(argv[1]) + ".bai");
    BamHeader h;
    h.first_record_voffset = strtoull(argv[2], nullptr, 10);
    h.reference_sequences.resize(strtoul(argv[3], nullptr, 10));
    for (const Shard& s : plan_shards(h, bai, (uint32_t)atoi(argv[4]), file.size())) {
      uint64_t lo = 0, hi = 0;
      if (!s.empty) shard_bytes(s, file.data(), file.size(), &lo, &hi);
      printf("%llu %llu %d %llu %llu", (unsigned long long)s.first_voffset, (unsigned long long)s.end_voffset, s.empty ? 1 : 0, (unsigned long long)lo, (unsigned long long)hi);
      for (uint32_t c : s.contigs) printf(" %u", c);
      printf("\n");
    }
  } catch (const std::exception& e) {
    printf("error: %s\n", e.what());
  }
  return 0;
}
