// C entry points over db_format.h for tests (ctypes): partitioner and suffix array.
#include <cstring>

#include "db_format.h"

extern "C" {

// part_of[k] = device of sequence k (LPT over `parts` devices)
int prib_lpt_partition(int n, const int *lens, int parts, int *part_of) {
  std::vector<std::string> fake((size_t)n);
  for (int k = 0; k < n; k++) fake[k].assign((size_t)lens[k], 'A');
  std::vector<std::vector<int>> part;
  prib::lpt_partition(fake, parts, part);
  for (size_t d = 0; d < part.size(); d++)
    for (int idx : part[d]) part_of[idx] = (int)d;
  return 0;
}

// the HOST checker of the GPU builder (the C-ABI prib_suffix_array lives in libpriblast_acc.so)
int prib_suffix_array_host(const unsigned char *text, int n, int *sa) {
  std::vector<int32_t> v;
  prib::build_suffix_array(text, n, v);
  std::memcpy(sa, v.data(), sizeof(int32_t) * v.size());
  return 0;
}
}
