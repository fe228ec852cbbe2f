// Drives the product's host-side facets exactly as ngs_b200/host/qc_command.cpp does after ngsq_finish
// (ingest -> summarize / setup -> ingest -> teardown -> aggregate -> Results::write), against the test
// double in fake_engine.cpp.  usage: host_results <refs.tsv (name \t length per line)> <out dir> <prefix>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../ngs_b200/host/facets.hpp"
#include "../../ngs_b200/host/genome.hpp"
#include "../../ngs_b200/host/results.hpp"

using namespace ngs;

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::vector<ReferenceSequence> refs;
  std::ifstream in(argv[1]);
  std::string name;
  uint32_t len;
  while (in >> name >> len) refs.push_back({name, len});
  auto genome = get_reference_genome("GRCh38_no_alt_AnalysisSet");
  FacetSet facets = get_qc_facets(*genome, std::nullopt);
  ngsq_engine* root = reinterpret_cast<ngsq_engine*>(1);  // opaque to the getters of the test double
  for (auto& f : facets.record_based) { f->ingest(root); f->summarize(); }
  for (auto& f : facets.sequence_based) f->ingest_global(root);
  for (uint32_t c = 0; c < refs.size(); ++c)
    for (auto& f : facets.sequence_based) {
      if (!f->supports_sequence_name(refs[c].name)) continue;
      f->setup(refs[c]);
      f->ingest(root, c, refs[c]);
      f->teardown(refs[c]);
    }
  Results results;
  for (auto& f : facets.record_based) f->aggregate(results);
  for (auto& f : facets.sequence_based) f->aggregate(results);
  results.write(argv[3], argv[2]);
  return 0;
}
