// Host side of the two "next" facets (ngs_b200/host/gene_model.hpp, facets.hpp, results.hpp) against the test double
// of the C ABI (fake_engine.cpp).
//   host_next_rows gff <file> <5 names>      prints "seq start stop class" per kept feature, or "error: <message>"
//   host_next_rows fasta <file>              prints "name length first-16-letters" per record
//   host_next_rows json                      ingests the integers of $NGSQ_FAKE_NEXT and prints the results JSON
//   host_next_rows vaf <fasta> <out> <name:len>...   EditsFacet with a VAF file: per-sequence ingest of $NGSQ_FAKE_VAF (edits.rs:317-340)
#include <iostream>

#include "../../ngs_b200/host/facets.hpp"

using namespace ngs;

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "";
  try {
    if (mode == "gff" && argc == 8) {
      FeatureNames names;
      for (int j = 0; j < 5; ++j) names.slot[j] = argv[3 + j];
      auto genome = get_reference_genome("GRCh38_no_alt_AnalysisSet");
      GenomicFeaturesFacet f = GenomicFeaturesFacet::try_from(argv[2], names, *genome);
      const auto sc = names.slot_class();
      std::cout << "slot_class";
      for (int j = 0; j < 5; ++j) std::cout << " " << (int)sc[j];
      std::cout << "\n";
      for (auto& kv : f.model())
        for (size_t i = 0; i < kv.second.start.size(); ++i)
          std::cout << kv.first << " " << kv.second.start[i] << " " << kv.second.stop[i] << " " << (int)kv.second.cls[i] << "\n";
      return 0;
    }
    if (mode == "fasta" && argc == 3) {
      for (auto& kv : read_fasta(slurp_maybe_gz(argv[2], "reference FASTA"))) std::cout << kv.first << " " << kv.second.size() << " " << kv.second.substr(0, 16) << "\n";
      return 0;
    }
    if (mode == "json") {
      ngsq_engine* root = reinterpret_cast<ngsq_engine*>(1);
      GenomicFeaturesFacet ff;
      ff.ingest(root);
      ff.summarize();
      EditsFacet ef;
      ef.ingest_global(root);
      Results r;
      ff.aggregate(r);
      ef.aggregate(r);
      std::cout << r.to_json_pretty();
      return 0;
    }
    if (mode == "vaf" && argc >= 5) {
      ngsq_engine* root = reinterpret_cast<ngsq_engine*>(1);
      EditsFacet ef = EditsFacet::try_from(argv[2], std::string(argv[3]));
      for (int k = 4; k < argc; ++k) {
        const std::string a = argv[k];
        ReferenceSequence seq;
        seq.name = a.substr(0, a.find(':'));
        seq.length = (uint32_t)std::stoul(a.substr(a.find(':') + 1));
        ef.setup(seq);
        ef.ingest(root, (uint32_t)(k - 4), seq);
        ef.teardown(seq);
      }
      Results r;
      ef.ingest_global(root);
      ef.aggregate(r);
      std::cout << "ok\n";
      return 0;
    }
  } catch (const std::exception& ex) {
    std::cout << "error: " << ex.what() << "\n";
    return 0;
  }
  return 2;
}
