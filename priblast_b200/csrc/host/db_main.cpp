// `pRIblast db` front-end on top of libpriblast_acc.so: same options and defaults as the reference
// (main.cpp:43-73, db_construction_parameters.cpp:32-78), same five output files.
// The per-sequence OpenMP/MPI loop of DbConstruction::CalculateAccessibility (db_construction.cpp:170-229)
// becomes ONE prib_acc_run call per GPU; the block/heap/dynamic distributors (-a) all map to the
// length-balanced partitioner over the GPUs of this box (db_format.h lpt_partition); output order is the
// FASTA order, i.e. what the reference emits with -np 1 and -a heap|block.
//
// Process model: one WORKER PROCESS per GPU (the reference: one MPI rank per node), forked before the first CUDA
// call and pinned to its device through CUDA_VISIBLE_DEVICES, so the CUDA start-up of eight B200s (0.55 s each,
// 5.4 s when one process opens all of them: profiles/r1/cuda_init_8gpu.txt) runs in parallel.  Every worker writes
// the records of its sequences straight into <db>.acc at their final offsets (pwrite; the record sizes are known
// from the lengths), so there is no gather step at all.  The parent builds <db>.seq/.ind/.nam/.bas meanwhile (suffix
// arrays on the last GPU).
#include <fcntl.h>
#include <getopt.h>
#include <sys/wait.h>
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
      "environment: PRIB_NUM_GPUS=n limits the number of GPUs (default: every visible one, one worker process each)\n");
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
  // record k of <db>.acc starts at rec_off[k] (raccess.cpp:447-481: 8 + 4 (2L - delta + 1) bytes per sequence)
  std::vector<int64_t> rec_off(seqs.size() + 1, 0);
  for (size_t k = 0; k < seqs.size(); k++) rec_off[k + 1] = rec_off[k] + prib_acc_record_bytes((int32_t)seqs[k].size(), delta);
  std::vector<pid_t> workers;
  std::vector<int> status_fd;
  std::vector<std::string> devs;
  if (!formats_only) {
    // the devices we may use: the caller's CUDA_VISIBLE_DEVICES list if there is one, else the /dev/nvidiaN nodes
    // (counted without touching CUDA: nothing in this process may initialise it before the fork)
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
    if (const char *e = std::getenv("PRIB_NUM_GPUS"))
      if (std::atoi(e) >= 1 && (size_t)std::atoi(e) < devs.size()) devs.resize((size_t)std::atoi(e));
    if (devs.empty()) return die("Error: no CUDA device available (there is no CPU path)");
    const int ngpu = (int)std::min(devs.size(), std::max<size_t>(seqs.size(), 1));
    devs.resize((size_t)ngpu);
    if (timer.on) std::fprintf(stderr, "[db] using %d GPU(s), one worker process each\n", ngpu);
    {  // <db>.acc at its final size; the workers fill it in place
      const int fd = open((db + ".acc").c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
      if (fd < 0 || ftruncate(fd, (off_t)rec_off[seqs.size()]) != 0) return die("Error: can't open " + db + ".acc");
      close(fd);
    }
    std::vector<std::vector<int>> part;
    lpt_partition(seqs, ngpu, part);
    std::fflush(nullptr);
    for (int d = 0; d < ngpu; d++) {
      int fds[2];
      if (pipe(fds) != 0) return die("Error: pipe failed");
      const pid_t pid = fork();
      if (pid < 0) return die("Error: fork failed");
      if (pid == 0) {  // ---- worker d: its own CUDA context on its own device
        close(fds[0]);
        setenv("CUDA_VISIBLE_DEVICES", devs[(size_t)d].c_str(), 1);
        const std::vector<int> &ids = part[(size_t)d];
        // the worker reports through the pipe as soon as its files are complete and exits afterwards: tearing down a
        // CUDA context with ~100 GB of DP state takes about a second that nobody has to wait for
        auto finish = [&](char status) {
          if (write(fds[1], &status, 1) != 1) _exit(2);
          close(fds[1]);
          _exit(status == 'k' ? 0 : 1);
        };
        auto fail_w = [&](const std::string &msg) {
          std::fprintf(stderr, "Error: GPU %s: %s\n", devs[(size_t)d].c_str(), msg.c_str());
          finish('e');
        };
        // <db>.seq / <db>.ind need only the sequences: the LAST worker (the lightest share under LPT ties) builds them
        // in a second thread of its process, suffix arrays on its GPU (one CUDA start-up per device)
        std::string err_si;
        std::future<bool> seq_ind;
        if (d == ngpu - 1)
          seq_ind = std::async(std::launch::async, [&]() {
            StageTimer t_si;
            const bool ok_si = write_seq_ind(db, seqs, prm, err_si, gpu_suffix_array);
            t_si.lap("  worker: .seq/.ind (SA on GPU, hash) [thread]");
            return ok_si;
          });
        auto join_si = [&]() {
          if (seq_ind.valid() && !seq_ind.get()) fail_w(err_si);
        };
        if (ids.empty()) {
          join_si();
          finish('k');
        }
        StageTimer wt;
        prib_acc_params ap;
        std::memset(&ap, 0, sizeof(ap));
        ap.maximal_span = prm.maximal_span;
        ap.min_accessible_length = delta;
        ap.device = 0;
        ap.mode = acc_mode;
        prib_ctx *ctx = nullptr;
        if (prib_acc_create(&ctx, &ap) != PRIB_OK) fail_w(prib_last_error());
        wt.lap("  worker: context (CUDA init, tables)");
        std::vector<const char *> sp(ids.size());
        std::vector<int32_t> sl(ids.size());
        std::vector<int64_t> ao(ids.size()), co(ids.size());
        int64_t total = 0;
        for (size_t k = 0; k < ids.size(); k++) {
          sp[k] = seqs[(size_t)ids[k]].data();
          sl[k] = (int32_t)seqs[(size_t)ids[k]].size();
          ao[k] = total;
          co[k] = total + sl[k];
          total += 2 * (int64_t)sl[k];
        }
        float *image = (float *)prib_host_alloc(sizeof(float) * (size_t)std::max<int64_t>(total, 1));
        if (!image) fail_w(prib_last_error());
        if (prib_acc_run(ctx, (int32_t)ids.size(), sp.data(), sl.data(), image, ao.data(), co.data()) != PRIB_OK)
          fail_w(prib_last_error());
        wt.lap("  worker: prib_acc_run");
        const int fd = open((db + ".acc").c_str(), O_WRONLY);
        if (fd < 0) fail_w("can't open " + db + ".acc");
        std::vector<char> rec;
        for (size_t k = 0; k < ids.size(); k++) {
          const int64_t bytes = rec_off[(size_t)ids[k] + 1] - rec_off[(size_t)ids[k]];
          rec.resize((size_t)bytes);
          prib_acc_write_record(image + ao[k], image + co[k], sl[k], delta, rec.data());
          if (pwrite(fd, rec.data(), (size_t)bytes, (off_t)rec_off[(size_t)ids[k]]) != (ssize_t)bytes)
            fail_w("short write on " + db + ".acc");
        }
        close(fd);
        wt.lap("  worker: records -> .acc");
        join_si();
        finish('k');
      }
      close(fds[1]);
      workers.push_back(pid);
      status_fd.push_back(fds[0]);
    }
  }
  bool ok = true;
  std::string err_si;
  if (formats_only && !write_seq_ind(db, seqs, prm, err_si, nullptr)) {  // test switch: host suffix arrays, no GPU
    ok = false;
    err = err_si;
  }
  if (ok && !write_nam(db, names, err)) ok = false;
  if (ok && !write_bas(db, prm, err)) ok = false;
  timer.lap(".nam/.bas");
  bool workers_ok = true;
  for (int fd : status_fd) {  // one status byte per worker; its process may still be tearing down afterwards
    char st = 0;
    if (read(fd, &st, 1) != 1 || st != 'k') workers_ok = false;
    close(fd);
  }
  for (pid_t pid : workers) waitpid(pid, nullptr, WNOHANG);
  timer.lap("accessibility (GPU workers) + .acc/.seq/.ind");
  if (!workers_ok) return die("Error: an accessibility worker failed (see above); the database is incomplete");
  if (!ok) return die(err);
  return 0;
}
