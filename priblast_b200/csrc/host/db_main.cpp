// `pRIblast db` front-end on top of libpriblast_acc.so: same options and defaults as the reference
// (main.cpp:43-73, db_construction_parameters.cpp:32-78), same five output files.
// The per-sequence OpenMP/MPI loop of DbConstruction::CalculateAccessibility (db_construction.cpp:170-229)
// becomes ONE prib_acc_run call per GPU; the block/heap/dynamic distributors (-a) all map to the
// length-balanced partitioner over the GPUs of this box (db_format.h lpt_partition); output order is the
// FASTA order, i.e. what the reference emits with -np 1 and -a heap|block.
#include <getopt.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/priblast_acc.h"
#include "db_format.h"

using namespace prib;

static int die(const std::string &msg) {
  std::fprintf(stderr, "%s\n", msg.c_str());
  return 1;
}

// <db>.ind: the suffix array of every page is built on GPU 0 (prib_suffix_array replaces sais(),
// db_construction.cpp:330-335)
static bool gpu_suffix_array(const uint8_t *text, int n, std::vector<int32_t> &sa, std::string &err) {
  sa.resize((size_t)std::max(n, 0));
  if (prib_suffix_array(text, n, sa.data(), 0) != PRIB_OK) {
    err = std::string("Error: suffix array: ") + prib_last_error();
    return false;
  }
  return true;
}

// PRIB_DB_TIMING=1: wall time of every stage on stderr
struct StageTimer {
  bool on = std::getenv("PRIB_DB_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char *what) {
    const auto now = std::chrono::steady_clock::now();
    if (on) std::fprintf(stderr, "[db] %-28s %9.3f s\n", what, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};

// GPUs of this box without touching CUDA: initialising the CUDA driver costs ~0.55 s + ~0.7 s per additional visible
// B200 (5.4 s with 8 on an NVSwitch box, profiles/r1/cuda_init_8gpu.txt), so the front-end decides how many GPUs a
// job is worth BEFORE the first CUDA call and hides the rest through CUDA_VISIBLE_DEVICES.
static int count_device_nodes() {
  int n = 0;
  for (int k = 0; k < 64; k++) {
    const std::string node = "/dev/nvidia" + std::to_string(k);
    if (FILE *f = std::fopen(node.c_str(), "rb")) {
      std::fclose(f);
      ++n;
    } else if (access(node.c_str(), F_OK) == 0) {
      ++n;
    }
  }
  return n;
}

// wall-time model: driver start-up 0.55 + 0.69 (n - 1) s, accessibility at 5.5e7 nt/s per GPU end to end
static int pick_gpus(double total_nt, int avail) {
  int best = 1;
  double best_t = 1e300;
  for (int n = 1; n <= avail; n++) {
    const double t = 0.55 + 0.69 * (n - 1) + total_nt / 5.5e7 / n;
    if (t < best_t) {
      best_t = t;
      best = n;
    }
  }
  return best;
}

static void usage() {
  std::printf(
      "pRIblast-b200 db: database construction with GPU accessibility\n"
      "usage: pRIblast_b200 db -i InputFastaFile -o OutputDbName [-r RepeatMaskingStyle] [-s LookupTableSize]\n"
      "                        [-w MaximalSpan] [-d MinAccessibleLength] [-c ChunkSize] [-a block|heap|dynamic]\n"
      "                        [-p TmpPath] [-m auto|fp64|exact]\n"
      "defaults: -r 0 -s 8 -w 70 -d 5 -a heap -c INT_MAX   (reference: main.cpp:43-73)\n"
      "-m (extension; also PRIB_ACC_MODE): auto = FP32 span-scaled engine with FP64 re-run (default; .acc within\n"
      "   1e-4 kcal/mol of the reference), fp64, exact = the reference's own arithmetic on the GPU (.acc, hence\n"
      "   the whole database and the ris output, byte-identical to the reference; ~50x slower than auto)\n"
      "environment: PRIB_NUM_GPUS=n fixes the number of GPUs (default: as many as the job is worth, see pick_gpus)\n");
}

int main(int argc, char *argv[]) {
  if (argc == 1 || std::strcmp(argv[1], "-h") == 0) {
    usage();
    return 0;
  }
  if (std::strcmp(argv[1], "db") != 0) {
    std::printf("usage: pRIblast_b200 [-h] db options   (the `ris` step stays the reference binary)\n");
    return 0;
  }
  std::string input, db, tmp_path, alg = "heap";
  std::string mode = std::getenv("PRIB_ACC_MODE") ? std::getenv("PRIB_ACC_MODE") : "auto";
  DbParams prm;
  int c;
  optind = 1;
  while ((c = getopt(argc - 1, argv + 1, "i:o:r:s:w:d:t:p:a:c:m:")) != -1) {
    switch (c) {
      case 'i': input = optarg; break;
      case 'o': db = optarg; break;
      case 'r': prm.repeat_flag = std::atoi(optarg); break;
      case 's': prm.hash_size = std::atoi(optarg); break;
      case 'w': prm.maximal_span = std::atoi(optarg); break;
      case 'd': prm.min_accessible_length = std::atoi(optarg); break;
      case 'p': tmp_path = optarg; break;  // accepted; there are no temp files any more
      case 'a': alg = optarg; break;
      case 'c': prm.chunk_size = std::atoi(optarg); break;
      case 'm': mode = optarg; break;
      default: return die("Error: invalid argument");  // incl. -t, as in the reference
    }
  }
  if (alg != "block" && alg != "heap" && alg != "dynamic") return die("Error: parallel algorithm not supported");
  if (mode != "auto" && mode != "fp64" && mode != "exact") return die("Error: -m must be auto, fp64 or exact");
  const int acc_mode = mode == "exact" ? 2 : mode == "fp64" ? 1 : 0;

  std::vector<std::string> names, seqs;
  std::string err;
  StageTimer timer;
  if (!read_fasta(input, names, seqs, err)) return die(err);
  timer.lap("read FASTA");
  if (db.empty()) return die("Error: -o option is required");                                   // raccess.hpp:42-45
  if (prm.min_accessible_length <= 1) return die("Error: -d option must be greater than 1");   // raccess.hpp:47-50
  if (prm.repeat_flag < 0 || prm.repeat_flag > 2) return die("Error: -r option must be 0, 1, or 2");
  const int delta = prm.min_accessible_length;
  for (size_t k = 0; k < seqs.size(); k++)
    if ((int)seqs[k].size() < delta)
      return die("Error: sequence " + names[k] + " is shorter than the minimum accessible length (-d): the "
                 "reference writes a corrupt record for it (raccess.cpp:449-450); refusing");

  // ---- accessibility on the GPUs ---------------------------------------------------------------
  const bool formats_only = std::getenv("PRIB_DB_FORMATS_ONLY") != nullptr;  // test switch: no .acc, no GPU
  std::vector<int64_t> acc_off(seqs.size()), cond_off(seqs.size());
  int64_t total = 0;
  for (size_t k = 0; k < seqs.size(); k++) {
    acc_off[k] = total;
    cond_off[k] = total + (int64_t)seqs[k].size();
    total += 2 * (int64_t)seqs[k].size();
  }
  int want = 0;
  if (!formats_only) {
    // how many GPUs is this job worth?  Decided before the first CUDA call (see count_device_nodes)
    if (const char *e = std::getenv("PRIB_NUM_GPUS")) want = std::max(1, std::atoi(e));
    // the devices we may use: the caller's CUDA_VISIBLE_DEVICES list if there is one, else the /dev/nvidiaN nodes
    std::vector<std::string> devs;
    if (const char *cv = std::getenv("CUDA_VISIBLE_DEVICES")) {
      std::string item;
      for (const char *q = cv;; ++q) {
        if (*q == ',' || *q == 0) {
          if (!item.empty()) devs.push_back(item);
          item.clear();
          if (*q == 0) break;
        } else {
          item += *q;
        }
      }
    } else {
      const int nodes = count_device_nodes();
      for (int k = 0; k < nodes; k++) devs.push_back(std::to_string(k));
    }
    if (devs.size() > 1) {
      const int n = std::min((int)devs.size(), want > 0 ? want : pick_gpus((double)total / 2, (int)devs.size()));
      if (n < (int)devs.size()) {
        std::string vis;
        for (int k = 0; k < n; k++) vis += (k ? "," : "") + devs[k];
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
      }
    }
  }
  // <db>.seq / <db>.ind need only the sequences: they are built (suffix arrays on GPU 0) while the GPUs work on
  // the accessibility
  std::string err_si;
  StageTimer t_si;
  std::future<bool> seq_ind = std::async(std::launch::async, [&]() {
    const bool ok = write_seq_ind(db, seqs, prm, err_si, formats_only ? nullptr : gpu_suffix_array);
    t_si.lap(".seq/.ind (SA on GPU, hash) [overlapped]");
    return ok;
  });
  float *image = nullptr;
  std::vector<std::thread> workers;
  if (!formats_only) {
    int ngpu = prib_device_count();
    if (want > 0) ngpu = std::min(ngpu, want);
    if (timer.on) std::fprintf(stderr, "[db] using %d GPU(s)\n", ngpu);
    if (ngpu <= 0) {
      seq_ind.wait();
      return die("Error: no CUDA device available (there is no CPU path)");
    }
    image = (float *)prib_host_alloc(sizeof(float) * (size_t)std::max<int64_t>(total, 1));
    if (!image) {
      seq_ind.wait();
      return die(std::string("Error: ") + prib_last_error());
    }
    timer.lap("pinned output image");
    std::vector<std::vector<int>> part;
    lpt_partition(seqs, ngpu, part);
    std::vector<std::string> errors(ngpu);
    std::vector<std::promise<void>> done(ngpu);  // results of GPU d are in the image (its context may still be tearing down)
    for (int d = 0; d < ngpu; d++) {
      workers.emplace_back([&, d]() {
        const std::vector<int> &ids = part[d];
        if (ids.empty()) {
          done[d].set_value();
          return;
        }
        prib_acc_params ap;
        std::memset(&ap, 0, sizeof(ap));
        ap.maximal_span = prm.maximal_span;
        ap.min_accessible_length = delta;
        ap.device = d;
        ap.mode = acc_mode;
        prib_ctx *ctx = nullptr;
        StageTimer wt;
        if (prib_acc_create(&ctx, &ap) != PRIB_OK) {
          errors[d] = prib_last_error();
          done[d].set_value();
          return;
        }
        wt.lap("  context (CUDA init, tables)");
        std::vector<const char *> sp(ids.size());
        std::vector<int32_t> sl(ids.size());
        std::vector<int64_t> ao(ids.size()), co(ids.size());
        for (size_t k = 0; k < ids.size(); k++) {
          sp[k] = seqs[ids[k]].data();
          sl[k] = (int32_t)seqs[ids[k]].size();
          ao[k] = acc_off[ids[k]];
          co[k] = cond_off[ids[k]];
        }
        if (prib_acc_run(ctx, (int32_t)ids.size(), sp.data(), sl.data(), image, ao.data(), co.data()) != PRIB_OK)
          errors[d] = prib_last_error();
        wt.lap("  prib_acc_run");
        done[d].set_value();
        prib_acc_destroy(ctx);  // freeing the DP state overlaps the file writing below
        wt.lap("  context teardown [overlapped]");
      });
    }
    for (int d = 0; d < ngpu; d++) done[d].get_future().wait();
    for (int d = 0; d < ngpu; d++)
      if (!errors[d].empty()) {
        for (auto &w : workers) w.join();
        seq_ind.wait();
        return die("Error: GPU " + std::to_string(d) + ": " + errors[d]);
      }
  }

  timer.lap("accessibility (GPU)");
  // ---- database files --------------------------------------------------------------------------
  bool ok = true;
  if (!formats_only && !write_acc(db, seqs, image, acc_off, cond_off, delta, err)) ok = false;
  timer.lap(".acc");
  if (ok && !write_nam(db, names, err)) ok = false;
  if (ok && !write_bas(db, prm, err)) ok = false;
  timer.lap(".nam/.bas");
  if (!seq_ind.get()) {
    ok = false;
    err = err_si;
  }
  timer.lap("wait for .seq/.ind");
  for (auto &w : workers) w.join();
  timer.lap("wait for context teardown");
  if (!ok) return die(err);
  if (image) prib_host_free(image);
  return 0;
}
