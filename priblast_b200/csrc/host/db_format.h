// Host glue of the `db` step: FASTA reading, base encoding, the paged suffix array + k-mer interval
// hash, and byte-compatible writers of <db>.bas/.nam/.acc/.seq/.ind (format: SURVEY §2.2).
// Re-written from the behaviour of the reference (citations per function); nothing here runs on the GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace prib {

// FASTA semantics of FastafileReader::ReadSeqs (fastafile_reader.cpp:76-133): the first line is a
// header; name = header minus '>'; sequence lines are concatenated verbatim after stripping one trailing
// CR/LF.  Returns false (and sets err) if the file cannot be opened.
bool read_fasta(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs,
                std::string &err);

// Encoder (encoder.hpp:36-78): sentinel 0, unknown 1, ACGU(T) 2..5; repeat_flag 1 keeps lower case as
// 6..9, 2 folds it.  Database sequences are stored REVERSED followed by the sentinel (encoder.cpp:27-36).
bool encode_reversed(const std::vector<std::string> &seqs, size_t first, size_t count, int repeat_flag,
                     std::vector<uint8_t> &out);

// Suffix array of a byte text (the reference calls sais(), sais.cpp:656; the suffix array of a text is
// unique, so any correct construction yields the same bytes).  Host prefix doubling with radix passes:
// TEST INFRASTRUCTURE (checker of prib_suffix_array, and the PRIB_DB_FORMATS_ONLY test switch); the
// `db` front-end builds the array on the GPU.
void build_suffix_array(const uint8_t *text, int n, std::vector<int32_t> &sa);

// Suffix-array builder used by write_seq_ind: returns false and sets err on failure.
typedef bool (*SaBuilder)(const uint8_t *text, int n, std::vector<int32_t> &sa, std::string &err);

// ConstructHashForShortSubstring + Search (db_construction.cpp:337-369, 438-500): for every k-mer over
// {2,3,4,5} of length 1..hash_size the suffix-array interval [start, end], with the reference's
// empty-interval convention (start = 1, end = 0).
void build_kmer_hash(const std::vector<uint8_t> &text, const std::vector<int32_t> &sa, int hash_size,
                     std::vector<std::vector<int32_t>> &start_hash, std::vector<std::vector<int32_t>> &end_hash);

struct DbParams {
  int hash_size = 8, repeat_flag = 0, maximal_span = 70, min_accessible_length = 5;  // db_construction_parameters.hpp:46-49
  int chunk_size = 2147483647;
};

// Writers (db_construction.cpp:371-436, 502-576).  acc/cond: per sequence, L floats each, FASTA order.
bool write_bas(const std::string &db, const DbParams &p, std::string &err);
bool write_nam(const std::string &db, const std::vector<std::string> &names, std::string &err);
bool write_acc(const std::string &db, const std::vector<std::string> &seqs, const float *image,
               const std::vector<int64_t> &acc_off, const std::vector<int64_t> &cond_off, int delta, std::string &err);
bool write_seq_ind(const std::string &db, const std::vector<std::string> &seqs, const DbParams &p, std::string &err,
                   SaBuilder sa_builder = nullptr /* nullptr: the host checker */);

// Length-balanced partition of sequences over `parts` devices: longest-processing-time greedy on the
// cost model c(L) = L (the DP is linear in L for fixed span), the same idea as the reference's heap
// distributor (fastafile_reader.cpp:248-314).  part[k] lists sequence indices, longest first.
void lpt_partition(const std::vector<std::string> &seqs, int parts, std::vector<std::vector<int>> &part);

}  // namespace prib
