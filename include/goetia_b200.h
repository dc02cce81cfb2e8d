/*
 * goetia_b200.h -- C ABI of the B200-native k-mer ingest backend (libgoetia_b200.so).
 *
 * goetia itself has no FFI for this path: the boundary in the reference is the C++ template
 * pair  dBG<StorageType, ShifterType>  (include/goetia/dbg.hh:39-41), reached from Python
 * through cppyy.  This header is the thin `extern "C"` layer that a GPU-backed StorageType /
 * ShifterType pair sits on (see INTEGRATION.md for the C++ adapter a goetia maintainer adds,
 * and goetia_b200/ for the Python mirror of the cppyy surface).  Every entry point cites the
 * reference member(s) it stands in for; paths are relative to the goetia source tree.
 *
 * Conventions
 *   - plain pointers and sizes only; handles are opaque; no CUDA or torch types.
 *   - functions returning int: 0 = ok, <0 = error (message via gt_last_error()).
 *     functions returning int64_t: >=0 = count (k-mers consumed), <0 = error.
 *     No exception crosses this boundary; the wrappers re-throw (reference error convention:
 *     GoetiaException family, goetia.hh:140-178).
 *   - one host thread drives one handle at a time (the reference dBG is not thread-safe
 *     either: kmeriterator.hh:64-76 re-bases the graph's own shifter).
 *   - "host" pointers are ordinary CPU memory (pinned memory makes the copies asynchronous);
 *     "_dev" entry points take device pointers already resident in HBM.  The library runs on
 *     its own non-blocking streams (or the ones given to gt_set_compute_stream): whatever
 *     produced those buffers must have completed, or be ordered before the call on that stream.
 *     d_bases must be 16-byte aligned (the packer reads it with 16 B loads), d_offsets 8-byte.
 *   - sequences travel as one concatenated byte buffer `bases` plus `offsets[n_reads+1]`
 *     (read r = bases[offsets[r] .. offsets[r+1])).  Per read the semantics are those of
 *     FastxParser<DNA_SIMPLE> + InserterProcessor (parsing/readers.hh:150-219,
 *     processors.hh:304-331): a/c/g/t are folded to upper case; a read holding any other byte
 *     is skipped (status bit GT_READ_INVALID); a read shorter than K contributes 0 k-mers
 *     (status bit GT_READ_SHORT) -- where the reference's dBG::insert_sequence would throw
 *     SequenceLengthException (kmeriterator.hh:57-59) and the processor would swallow it.
 *   - all hash / bin arithmetic is unsigned 64-bit and bit-exact with the reference.
 */
#ifndef GOETIA_B200_H
#define GOETIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT_ABI_VERSION 1

/* StorageType selector: BitStorage (storage/bitstorage.hh:93), ByteStorage
 * (storage/bytestorage.hh:96), NibbleStorage (storage/nibblestorage.hh:89). */
enum { GT_STORAGE_BIT = 0, GT_STORAGE_BYTE = 1, GT_STORAGE_NIBBLE = 2 };
/* ShifterType selector: FwdLemireShifter / CanLemireShifter (hashing/hashshifter.hh:204-205). */
enum { GT_SHIFTER_FWD = 0, GT_SHIFTER_CAN = 1 };
/* How order-dependent outputs (is_new, n_unique) are produced -- SURVEY.md section 8a
 * "sequential-equivalence rule".  Final table bytes are identical in every mode.
 *   GT_MODE_BLIND : fire-and-forget updates; n_unique is NOT maintained (stats report it
 *                   as of the last tracked batch); fastest.
 *   GT_MODE_FAST  : a k-mer is credited as new when one of ITS atomic updates found the
 *                   slot empty ("atomic winner"); equals the serial count except when two
 *                   different k-mers of one batch collide on a fresh slot.
 *   GT_MODE_EXACT : serial-order semantics (first toucher in read/k-mer order), identical
 *                   to the reference's one-at-a-time loop. */
enum { GT_MODE_BLIND = 0, GT_MODE_FAST = 1, GT_MODE_EXACT = 2 };
/* per-read status bits */
enum { GT_READ_OK = 0, GT_READ_SHORT = 1, GT_READ_INVALID = 2 };

typedef struct gt_storage gt_storage; /* device-resident tables of one StorageType */
typedef struct gt_batch gt_batch;     /* device-resident 2-bit packed read batch     */
typedef struct gt_sketch gt_sketch;   /* device-resident scaled / bottom-k MinHash   */

/* ---- library ------------------------------------------------------------------------ */
int gt_abi_version(void);
/* Select the CUDA device this process drives (one process per GPU).  Must precede any other
 * call; repeated calls with the same device are no-ops. */
int gt_init(int device);
int gt_device_count(void);
/* Last error message of the calling thread (sourmash.hpp:40-65 polls a last-error code the
 * same way). */
const char* gt_last_error(void);
/* Blocks until all work queued by this library has finished. */
int gt_synchronize(void);

/* ---- table sizing ------------------------------------------------------------------- */
/* get_n_primes_near_x (storage/storage.hh:166-190): the n largest primes strictly below x,
 * descending.  Returns how many were written (can be < n for tiny x). */
int gt_primes_near(uint32_t n, uint64_t x, uint64_t* out);

/* ---- StorageType -------------------------------------------------------------------- */
/* Storage(const std::vector<uint64_t>& tablesizes): bitstorage.hh:113-121,
 * bytestorage.hh (ctor), nibblestorage.hh:140-149.  Tables are zero-initialised in HBM. */
gt_storage* gt_storage_create(int kind, const uint64_t* tablesizes, int n_tables);
void gt_storage_destroy(gt_storage* st);
/* reset(): zero all tables and both counters (bitstorage.cc reset, nibblestorage.hh:179+). */
int gt_storage_reset(gt_storage* st);
int gt_storage_kind(const gt_storage* st);
int gt_storage_n_tables(const gt_storage* st);             /* n_tables()      */
int gt_storage_tablesizes(const gt_storage* st, uint64_t* out); /* get_tablesizes() */
/* Bytes of table i exactly as the reference allocates them: size/8+1 (bitstorage.hh:143-156),
 * size (bytestorage.hh:125-134), size/2+1 (nibblestorage.hh:166-177). */
uint64_t gt_storage_table_bytes(const gt_storage* st, int i);
/* get_raw_tables() (storage.hh:126): copy table i to host memory, byte-identical to the
 * reference's table after the same inserts. */
int gt_storage_download_table(gt_storage* st, int i, uint8_t* host_dst);
/* load(): replace table i from host bytes (same layout); counters are NOT touched. */
int gt_storage_upload_table(gt_storage* st, int i, const uint8_t* host_src);
/* n_unique_kmers() / n_occupied() (storage.hh:119-120).  n_occupied is recomputed from
 * table 0 (number of non-zero slots), which is what the reference's counter equals. */
int gt_storage_stats(gt_storage* st, uint64_t* n_unique, uint64_t* n_occupied);
int gt_storage_set_n_unique(gt_storage* st, uint64_t n_unique);
/* Position-weighted checksum of table i, computed in HBM: sum over the table's little-endian 32-bit
 * words of word * weight(global word index) mod 2^64, weight(j) = fmix64(j + 0x9e3779b97f4a7c15) | 1
 * (fmix64 = the MurmurHash3 finaliser).  Linear, so the parts of a sharded storage sum to the checksum
 * of the reference's whole table; lets multi-GB tables be compared without moving them. */
int gt_storage_checksum(gt_storage* st, int i, uint64_t* out);
/* BitStorage::update_from (bitstorage.cc:103-137): dst |= src, same table sizes. */
int gt_storage_update_from(gt_storage* dst, const gt_storage* src);
/* Raw device pointer of table i (for peer mapping / torch interop). */
void* gt_storage_device_table(gt_storage* st, int i);

/* Storage::insert / query / insert_and_query over a vector of hash values
 * (bitstorage.hh:195-219, bitstorage.cc:78-100, bytestorage.cc:60-150,
 * nibblestorage.cc:60-130).  is_new (uint8 per hash) may be NULL.  Host pointers. */
int gt_insert_hashes(gt_storage* st, const uint64_t* hashes, uint64_t n, int mode, uint8_t* is_new);
int gt_query_hashes(gt_storage* st, const uint64_t* hashes, uint64_t n, int16_t* counts);

/* ---- ShifterType -------------------------------------------------------------------- */
/* KmerIterator<Shifter> over every read (hashing/kmeriterator.hh:64-123).  fw[] (and rc[]
 * for GT_SHIFTER_CAN; may be NULL for FWD) receive one value per k-mer, reads back to back
 * (read r's k-mers start at the running sum of max(0, len-K+1) over valid earlier reads).
 * status (uint8 per read) may be NULL.  Returns the total number of k-mers. */
int64_t gt_hash_sequences(int shifter, int K, const char* bases, const uint64_t* offsets,
                          uint64_t n_reads, uint64_t* fw, uint64_t* rc, uint8_t* status);

/* ---- dBG<Storage, Shifter> batch members (host buffers) ------------------------------- */
/* dBG::insert_sequence over a batch (dbg.hh:296-305; with n_new: :307-318).  Returns the
 * total k-mers consumed (sum of len-K+1).  n_new_per_read (uint64 per read) needs
 * GT_MODE_EXACT and may be NULL.  status may be NULL. */
int64_t gt_insert_sequences(gt_storage* st, int shifter, int K, const char* bases,
                            const uint64_t* offsets, uint64_t n_reads, int mode,
                            uint64_t* n_new_per_read, uint8_t* status);
/* dBG::query_sequence over a batch (dbg.hh:349-362): counts laid out like gt_hash_sequences. */
int64_t gt_query_sequences(gt_storage* st, int shifter, int K, const char* bases,
                           const uint64_t* offsets, uint64_t n_reads, int16_t* counts,
                           uint8_t* status);
/* DiginormFilter::median_count_at_least per read (diginorm.hh:35-68): pass[r] = 1 iff at
 * least unsigned(0.5 + float(n_kmers)/2) k-mers of read r have count >= cutoff.  Skipped
 * reads get 0. */
int64_t gt_median_count_at_least(gt_storage* st, int shifter, int K, const char* bases,
                                 const uint64_t* offsets, uint64_t n_reads, uint32_t cutoff,
                                 uint8_t* pass, uint8_t* status);

/* gt_median_count_at_least for reads already resident in HBM (ASCII d_bases, 16-byte aligned; uint64 d_offsets starting
 * at 0): pack + walk + decide queued on the compute stream with no host wait; d_pass (device, uint8 per read) receives
 * the decisions, d_kmer_total (device uint64, may be NULL) is incremented by the k-mers judged.  The C2 workload's
 * "per-read median-count query" (diginorm.hh:35-68 over dbg.hh:349-362). */
int gt_median_count_at_least_dev(gt_storage* st, int shifter, int K, const void* d_bases, const void* d_offsets,
                                 uint64_t n_reads, uint64_t n_bases, uint32_t cutoff, void* d_pass, void* d_kmer_total);

/* DiginormFilter::Filter::filter_sequence over a batch (diginorm.hh:111-119; FilterProcessor,
 * processors.hh:389-417), batch-synchronous: every read of the call is judged against the table state
 * at the start of the call, then the reads that passed (median count below cutoff) are inserted.  A call
 * of one read is the reference's serial filter.  keep[r] = 1 for passing reads, *n_kept = their number;
 * returns the k-mers of all judged reads (filter_sequence's second tuple member, summed). */
int64_t gt_diginorm_sequences(gt_storage* st, int shifter, int K, const char* bases,
                              const uint64_t* offsets, uint64_t n_reads, uint32_t cutoff, uint8_t* keep,
                              uint64_t* n_kept);

/* The same filter with the REFERENCE'S SERIAL semantics (diginorm.hh:111-119 judges and inserts read by read, so a read
 * sees every earlier kept read of the batch): rounds of judge / claim / verify / insert that decide, in parallel, exactly
 * what the one-at-a-time loop decides (see capi.cu).  Same arguments and return values.  This is what FilterProcessor
 * uses by default; gt_diginorm_sequences (batch-synchronous, one round) is the faster opt-in. */
int64_t gt_diginorm_sequences_serial(gt_storage* st, int shifter, int K, const char* bases,
                                     const uint64_t* offsets, uint64_t n_reads, uint32_t cutoff, uint8_t* keep,
                                     uint64_t* n_kept);
/* dBG::insert_and_query_sequence over a batch (dbg.hh:327-340; StreamingSolidFilter's input, solidifier.hh:58-76): every
 * k-mer is inserted and its count AFTER its own insert returned (Storage::insert_and_query: bitstorage.cc:78-84 always 1;
 * bytestorage.cc:142-150, nibblestorage.cc:102-109), with the serial semantics of the reference over the whole batch --
 * count_j = min_i min(max, pre_i + rank_j(i, bin) + 1), SURVEY.md section 8a.  counts laid out like gt_query_sequences. */
int64_t gt_insert_and_query_sequences(gt_storage* st, int shifter, int K, const char* bases,
                                      const uint64_t* offsets, uint64_t n_reads, int16_t* counts, uint8_t* status);

/* ---- device-resident batches (the parsing-to-device pipeline's product) --------------- */
/* Upload + validate + 2-bit pack a batch (A=0 C=1 G=2 T=3; flat base p at bits 2*(p%32) of
 * 64-bit word p/32).  The batch stays in HBM until destroyed and can be inserted / queried
 * any number of times. */
gt_batch* gt_batch_pack(const char* bases, const uint64_t* offsets, uint64_t n_reads);
/* Same, but from device-resident ASCII (d_bases, d_offsets are device pointers). */
gt_batch* gt_batch_pack_dev(const void* d_bases, const void* d_offsets, uint64_t n_reads,
                            uint64_t n_bases);
void gt_batch_destroy(gt_batch* b);
uint64_t gt_batch_n_reads(const gt_batch* b);
uint64_t gt_batch_n_bases(const gt_batch* b);
/* total k-mers a K-mer walk of this batch consumes */
int64_t gt_batch_n_kmers(gt_batch* b, int K);
int gt_batch_status(gt_batch* b, int K, uint8_t* status);
/* Queue insert / query of a resident batch on `stream` (0 = library stream); asynchronous. */
int64_t gt_insert_batch(gt_storage* st, int shifter, int K, gt_batch* b, int mode, void* stream);

/* dBG::insert_sequence over a batch whose ASCII bases and offsets (uint64, offsets[0] == 0) are
 * already resident in HBM.  Queues pack + hash + insert on the library's compute stream and
 * returns the k-mers consumed (one 8-byte read-back). */
int64_t gt_insert_sequences_dev(gt_storage* st, int shifter, int K, const void* d_bases,
                                const void* d_offsets, uint64_t n_reads, uint64_t n_bases, int mode);

/* The same without any host wait: everything is queued on the compute stream and the call returns at
 * once (0 / <0).  d_kmer_total (device uint64, may be NULL): the k-mers consumed are ADDED to it on
 * the stream.  The caller keeps d_bases / d_offsets alive until the stream has passed the call. */
int gt_insert_sequences_dev_async(gt_storage* st, int shifter, int K, const void* d_bases,
                                  const void* d_offsets, uint64_t n_reads, uint64_t n_bases, int mode,
                                  void* d_kmer_total);

/* ---- 2-bit packed host batches: the parsing-to-device pipeline's product --------------------------------------- */
/* Validate + fold case + 2-bit pack a batch ON THE HOST (what the FASTX front end's parser threads do, so that only
 * 0.25 B/base cross PCIe): words[ceil(n_bases/32)] in the device layout (base p of the batch at bits 2*(p%32) of word
 * p/32, counted from offsets[0]; A=0 C=1 G=2 T=3), flags[n_reads] = GT_READ_INVALID for reads holding a byte outside
 * ACGTacgt (FastxParser would skip them: parsing/readers.hh:162-171).  n_threads <= 0: all cores. */
int gt_pack_reads_host(const char* bases, const uint64_t* offsets, uint64_t n_reads, uint64_t* words, uint8_t* flags,
                       int n_threads);
/* dBG::insert_sequence (dbg.hh:296-305) over a packed host batch (pinned memory makes the copies asynchronous):
 * chunked H2D of words / offsets / flags on two streams, no pack kernel.  GT_MODE_BLIND or GT_MODE_FAST.  Returns the
 * k-mers consumed. */
int64_t gt_insert_sequences_packed(gt_storage* st, int shifter, int K, const uint64_t* words, const uint64_t* offsets,
                                   const uint8_t* flags, uint64_t n_reads, int mode);
/* The same for a packed batch already in HBM (d_words 16-byte aligned, bit 0 of word 0 = base 0; d_offsets uint64
 * starting at 0; d_flags uint8 per read).  n_words_alloc >= ceil(n_bases/32) + 1 words must be readable.  Queued on the
 * compute stream with no host wait; d_kmer_total (device uint64, may be NULL) is incremented by the k-mers consumed. */
int gt_insert_packed_dev_async(gt_storage* st, int shifter, int K, const void* d_words, uint64_t n_words_alloc,
                               const void* d_offsets, const void* d_flags, uint64_t n_reads, uint64_t n_bases, int mode,
                               void* d_kmer_total);

/* ---- write-combining of blind inserts --------------------------------------------------- */
/* GT_MODE_BLIND inserts into tables larger than L2 are not applied one by one: each update
 * is appended to the bucket of its table slice and applied slice by slice (DESIGN.md,
 * "write-combining insert").  Every call that reads a table (query, stats, download, save,
 * update_from, tracked inserts) applies what is pending first, so the semantics of
 * Storage::insert (bitstorage.hh:195-219 etc.) are unchanged; gt_storage_flush forces it. */
int gt_storage_flush(gt_storage* st);
/* Queue the apply of what is pending on the compute stream without waiting (gt_storage_flush
 * = this + wait).  On a sharded storage this is the step after the bucket exchange. */
int gt_storage_apply(gt_storage* st);
/* diagnostics: info[8] = {store built, n_buckets, log2(slots per slice), k-mer budget between
 * flushes, entries allocated, pending k-mers (upper bound), updates that overflowed a bucket
 * and were applied directly, apply-grid size} */
int gt_storage_pending_info(gt_storage* st, uint64_t* info);
/* number of CUDA kernels this library has launched so far in this process */
uint64_t gt_launch_count(void);
/* Harness helper (no counterpart in the reference; SURVEY.md section 8d asks for synthetic reads that the CPU
 * oracle and every GPU rank see byte-identically): write n_bases ASCII bytes of the counter-based synthetic
 * base stream `seed`, starting at GLOBAL base index first_base, to device memory (16-byte aligned), asynchronously
 * on the compute stream.  base(i) = "ACGT"[(splitmix64(seed + (i/32 + 1) * 0x9E3779B97F4A7C15) >> 2*(i%32)) & 3];
 * goetia_b200/synth.py is the numpy twin. */
int gt_synth_bases_dev(void* d_out, uint64_t n_bases, uint64_t seed, uint64_t first_base);
/* Harness helper: the random-access roofline of this device, measured now -- operations per second of independent
 * random 32-bit RED.OR (what = 0; SURVEY.md section 8d's R_rand for inserts) or random 32-byte sector loads (what = 1;
 * the query roofline) over a scratch footprint of footprint_bytes; device-timed, best of 3.  < 0 on error. */
double gt_probe_random(int what, uint64_t footprint_bytes, uint64_t n_ops);
/* Device timing for harnesses (the library launches on its own streams, which events of
 * another runtime's stream do not see).  gt_timer_record puts a CUDA event on the compute
 * stream (ordered after everything queued so far); gt_timer_elapsed_ms waits for both. */
int gt_timer_record(int slot);                       /* slot 0..7 */
double gt_timer_elapsed_ms(int from_slot, int to_slot);
/* Per-kernel device time of the insert kernels, from CUDA event pairs around each launch:
 * ms3/n3 = {k_bucket, k_apply, k_walk} accumulated since gt_profile_enable(1). */
int gt_profile_enable(int on);
int gt_profile_get(double* ms3, uint64_t* n3);
/* The same with the apply split up: ms/n[0..count) = {k_bucket, apply (whole), k_walk, k_rebucket, k_apply_win};
 * returns the number of kinds the library tracks. */
int gt_profile_get_detail(double* ms, uint64_t* n, int count);

/* ---- sharded storage: one process per GPU (no counterpart in the reference) -------------- */
/* Table t is cut into slices of 2^shift slots; rank r of `world` holds a contiguous run of
 * slices of every table, so concatenating the ranks' parts in rank order gives exactly the
 * single-GPU / reference table.  Every rank hashes its own reads, buckets the updates by slice
 * (k_bucket), ships the buckets of foreign slices to their owners, and applies what it
 * receives (k_apply).  The exchange itself belongs to the caller (goetia_b200/shard.py does it
 * with NCCL all-to-all); these entry points define the plan and the buffer layout.
 *
 * gt_shard_plan: host arithmetic only (no GPU needed).  shift_nb[2] = {log2(slots per slice),
 * number of buckets}; per bucket: table, owner rank, first slot, slots, capacity in entries
 * per producer and round (arrays of >= 1024 elements); own_lo/own_hi[world * n_tables] = slot
 * range of table t on rank r at [r * n_tables + t].  slice_log2_bytes <= 0 selects the default (32 MB of
 * table, 64 MB for the counting storages). */
int gt_shard_plan(int kind, const uint64_t* tablesizes, int n_tables, int world, uint64_t budget_kmers,
                  int slice_log2_bytes, int32_t* shift_nb, int32_t* table, int32_t* owner, uint64_t* slot0,
                  uint64_t* slots, uint32_t* cap, uint64_t* own_lo, uint64_t* own_hi);
/* This rank's part of a storage sharded over `world` ranks.  budget_kmers = the most k-mers
 * one rank buckets between two exchanges.  gt_storage_download_table returns the local part
 * (gt_storage_table_bytes bytes, laid out as that part of the reference's table). */
gt_storage* gt_storage_create_sharded(int kind, const uint64_t* tablesizes, int n_tables, int rank, int world,
                                      uint64_t budget_kmers, int slice_log2_bytes);
int gt_storage_local_range(const gt_storage* st, int i, uint64_t* lo, uint64_t* hi);
/* Device buffers of the exchange (caller-owned, e.g. torch tensors), all uint32:
 *   outbox    [sum of all regions]  regions of the peers in rank order (skipping this rank),
 *                                   then this rank's own region; a region = the buckets its
 *                                   owner holds, in bucket order, `cap` entries each
 *   inbox     [(world-1) * R_me]    one copy of this rank's region per peer, in rank order
 *   fill_send [n_buckets]           entry count of every bucket produced here (bucket order)
 *   fill_recv [world * n_owned]     [q][j] = count rank q produced for this rank's j-th bucket */
int gt_storage_attach_exchange(gt_storage* st, int which, void* outbox, void* inbox, void* fill_send,
                               void* fill_recv);
/* Two buffer sets (which = 0, 1) may be attached so that the exchange of one round overlaps
 * the hashing of the next: gt_storage_select_store picks the set the following inserts bucket
 * into, gt_storage_apply_store queues k_apply for one set on the apply stream (the caller
 * orders it after that set's exchange and before the set is bucketed into again). */
int gt_storage_select_store(gt_storage* st, int which);
int gt_storage_apply_store(gt_storage* st, int which);
/* Peer-memory transport (the fused hash + exchange path): instead of bucketing into a local
 * outbox that a collective then ships, k_bucket stores every run of entries straight into the
 * inbox of the slice's owner over NVLink (plain 16 B stores to peer memory mapped with CUDA IPC),
 * so the transfer overlaps the hashing run by run and no bulk all-to-all is left -- only the
 * per-bucket fill counts travel by collective, which is also the "all producers are done" signal.
 *   inbox of rank q (one per buffer set) = world regions of R_q entries, region p written by rank p
 *   (R_q = sum of the capacities of the buckets q holds; bucket b sits at its offset inside R_q).
 * gt_peer_alloc: zeroed device memory that can be exported; gt_peer_export: 64-byte handle to
 * send to the peers; gt_peer_open: map a peer's allocation into this process (enables peer access);
 * gt_peer_close / gt_peer_free undo them. */
void* gt_peer_alloc(uint64_t bytes);
int gt_peer_free(void* ptr);
int gt_peer_export(void* ptr, uint8_t handle[64]);
void* gt_peer_open(const uint8_t handle[64]);
int gt_peer_close(void* ptr);
/* Bytes of rank `rank`'s inbox (one per buffer set): the world bucket regions followed by world
 * overflow lists (GT_OVF_RECORDS, default 2^18, records of 8 bytes each).  An update whose bucket is
 * full (heavily duplicated k-mers) cannot be applied by the producer when the slice is foreign; it is
 * posted to the producer's overflow list in the owner's inbox as a full (table, slot) record and
 * applied there, so correctness does not depend on bucket capacities here either. */
uint64_t gt_storage_inbox_bytes(const gt_storage* st, int rank);
/* The same layout as host arithmetic (no GPU needed): region_entries[world] = R_q, in_region[n_buckets] =
 * entry offset of bucket b inside a region of its owner, ovf_offset_bytes[world] = where the overflow lists
 * start in rank q's inbox, inbox_bytes[world].  Rank p's entries for bucket b (owner q) land at entry
 * p * R_q + in_region[b] of q's inbox; its overflow list at ovf_offset_bytes[q] + p * GT_OVF_RECORDS * 8.
 * Returns the number of buckets. */
int gt_shard_peer_layout(int kind, const uint64_t* tablesizes, int n_tables, int world,
                         uint64_t budget_kmers, int slice_log2_bytes, uint64_t* region_entries,
                         uint64_t* in_region, uint64_t* ovf_offset_bytes, uint64_t* inbox_bytes);
/* inbox_of_rank[world]: device pointers valid in THIS process (entry `rank` = this rank's own
 * inbox of gt_storage_inbox_bytes bytes from gt_peer_alloc, the others from gt_peer_open).
 * fill_send [n_buckets + world]: bucket cursors, then the overflow-list cursors per owner rank;
 * fill_recv [world * (n_owned + 1)]: [q] = what rank q produced for this rank's n_owned buckets, then
 * its overflow count.  The caller moves fill_send -> fill_recv (owner-major) after k_bucket, e.g. with
 * one small all-to-all. */
int gt_storage_attach_peers(gt_storage* st, int which, void* const* inbox_of_rank, void* fill_send,
                            void* fill_recv);
/* Staged peer transport (copy engines instead of SM stores; the default of goetia_b200/shard.py): the inboxes are
 * laid out as above, but k_bucket writes what it produces for a FOREIGN owner q into a local staging area
 * stage_of_rank[q] (gt_storage_stage_bytes(st, q) bytes, 16-byte aligned: R_q entries, padded to 16 bytes, then this
 * rank's overflow list for q; entry `rank` of the array is ignored), and the caller ships the area into q's inbox --
 * the R_q entries to entry rank * R_q, the list to ovf_offset_bytes[q] + rank * GT_OVF_RECORDS * 8 -- with
 * gt_peer_copy_async on a copy stream: device-to-device copies over NVLink that run on the copy engines while the
 * SMs hash the next round.  This rank's own buckets go straight into region `rank` of own_inbox.  fill_send /
 * fill_recv as for gt_storage_attach_peers. */
uint64_t gt_storage_stage_bytes(const gt_storage* st, int rank);
int gt_storage_attach_staged(gt_storage* st, int which, void* own_inbox, void* const* stage_of_rank, void* fill_send,
                             void* fill_recv);
/* The general form: per foreign owner q, where k_bucket writes this rank's entries (region_of_rank[q]: R_q entries,
 * 16-byte aligned) and overflow records (ovf_of_rank[q]: GT_OVF_RECORDS records) for q -- a local staging area the
 * caller ships, or the final place in q's inbox mapped with gt_peer_open (then the stores cross NVLink inside k_bucket,
 * as with gt_storage_attach_peers).  goetia_b200/shard.py mixes the two per peer so that the SMs' stores and the copy
 * engines share the NVLink load.  Entry `rank` of both arrays is ignored. */
int gt_storage_attach_areas(gt_storage* st, int which, void* own_inbox, void* const* region_of_rank,
                            void* const* ovf_of_rank, void* fill_send, void* fill_recv);
/* cudaMemcpyAsync(dst, src, bytes) on `stream` (a cudaStream_t); dst / src: local device memory or a peer's
 * allocation mapped with gt_peer_open. */
int gt_peer_copy_async(void* dst, const void* src, uint64_t bytes, void* stream);
/* Upper bound of the k-mers in the NEXT device-resident batch inserted into `st` (gt_insert_sequences_dev[_async],
 * gt_insert_packed_dev_async), used once.  Without it the library assumes one k-mer per base (it cannot read the
 * device-resident offsets without a round trip), so a sharded storage's budget_kmers -- which sizes every exchange
 * buffer -- would have to be given in bases; callers with equal-length reads know the exact count. */
int gt_storage_hint_kmers(gt_storage* st, uint64_t n_kmers_upper);
/* Storage::query restricted to the slots this rank holds (routed queries on a sharded storage):
 * counts[i] = AND / min over the tables' slots of hashes[i] that fall in this rank's ranges, the
 * neutral element (1, 255, 15) where none does; the AND / min over all ranks' answers (an
 * all-reduce MIN) is exactly Storage::query (bitstorage.cc:87-100, bytestorage.cc:116-139,
 * nibblestorage.cc:112-130).  d_hashes / d_counts are device pointers; asynchronous on the
 * compute stream. */
int gt_query_hashes_local_dev(gt_storage* st, const void* d_hashes, uint64_t n, void* d_counts);
/* Owner-routed requests ("owners answer with the bit / count, 1 B per message"; the n_unique
 * reverse route of tracked inserts).  A request is (table << 59 | global slot), 8 bytes.
 *   gt_shard_route_hashes_dev  for n hash values: d_req[n * n_tables] requests and d_owner[n * n_tables] (int32) = the rank
 *                              holding each slot; the caller ships the requests to their owners (all-to-all);
 *   gt_shard_answer_dev        owner side of a query: d_answers[i] (uint8) = value of the slot of request i; the source
 *                              takes the AND / min over a hash's n_tables answers = Storage::query;
 *   gt_shard_insert_requests_dev  owner side of an insert: applies the requests; with d_ord (uint32 serial ordinals given by
 *                              the sources) d_first[i] = 1 iff request i is the first toucher of a slot that was empty
 *                              (GT_MODE_EXACT's rule); the source ORs a hash's flags = Storage::insert's return value.
 * All device pointers; asynchronous on the compute stream. */
int gt_shard_route_hashes_dev(gt_storage* st, const void* d_hashes, uint64_t n, void* d_req, void* d_owner);
int gt_shard_answer_dev(gt_storage* st, const void* d_req, uint64_t n_req, void* d_answers);
int gt_shard_insert_requests_dev(gt_storage* st, const void* d_req, const void* d_ord, uint64_t n_req, void* d_first);
/* KmerIterator over reads resident in HBM (ASCII d_bases, uint64 d_offsets from 0): d_values (uint64, capacity >= n_bases)
 * receives hash_type::value() of every k-mer, reads back to back.  Returns the number of k-mers. */
int64_t gt_hash_values_dev(int shifter, int K, const void* d_bases, const void* d_offsets, uint64_t n_reads,
                           uint64_t n_bases, void* d_values);
/* Run the library's kernels on caller-owned CUDA streams (NULL = the library's own), so that
 * they order with the caller's collectives: the compute stream carries pack / hash / bucket,
 * the apply stream k_apply. */
int gt_set_compute_stream(void* stream);
int gt_set_apply_stream(void* stream);

/* ---- FASTX front end: FastxParser<DNA_SIMPLE> + FileProcessor (the parsing-to-device pipeline) ---- */
/* FastxParser(infile, strict, min_length) (parsing/readers.hh:120-122, readers.cc:15-32): kseq
 * record grammar (parsing/kseq.h:174-214: FASTA / FASTQ, multi-line records, gz-transparent).  A reader
 * thread inflates ahead of the parser.  NULL + gt_last_error() if the file cannot be opened. */
typedef struct gt_fastx gt_fastx;
gt_fastx* gt_fastx_open(const char* path, int strict, uint32_t min_length);
void gt_fastx_close(gt_fastx* fx);
/* FastxParser::next (readers.hh:150-219).  1 = a record (pointers into the parser, valid until the next
 * call; the sequence is validated and upper-cased as DNA_SIMPLE::validate does); 0 = std::nullopt (the
 * record held a foreign symbol or was shorter than min_length and was skipped, or the end of the file
 * was reached -- see gt_fastx_stats); -2 = InvalidRead (sequence and quality lengths differ; parsing can
 * continue); -3 = GoetiaFileException (stream error); -4 = InvalidCharacterException (strict parsers:
 * a foreign symbol is an error instead of a skip); -5 = NoMoreReadsAvailable. */
int gt_fastx_next_record(gt_fastx* fx, const char** name, uint64_t* name_len, const char** seq,
                         uint64_t* seq_len, const char** qual, uint64_t* qual_len);
/* The same, batched for the device pipeline: appends the sequences of the next kept records to `bases`
 * (at most max_bases bytes) and their ends to offsets[1..n] (offsets[0] = 0, at most max_reads records).
 * Returns n (0 = end of file), or an error code as above. */
int64_t gt_fastx_next_batch(gt_fastx* fx, char* bases, uint64_t max_bases, uint64_t* offsets,
                            uint64_t max_reads);
/* The same in the device pipeline's own transfer format: the sequences 2-bit packed (base p of the batch at bits
 * 2*(p%32) of word p/32, A=0 C=1 G=2 T=3) -- what gt_insert_sequences_packed takes.  For uncompressed files the
 * parser's worker threads pack while they parse and whole verified pieces are handed out; every piece starts on a
 * word boundary and the gap behind it is covered by a phantom read flagged GT_READ_INVALID, so the return value n
 * (entries of offsets[1..n] and flags[0..n)) can exceed *n_real, the number of records.  words: at least
 * max_words (>= 2) u64; offsets: max_reads + 1; flags: max_reads.  0 = end of file, < 0 = error as above. */
int64_t gt_fastx_next_packed_batch(gt_fastx* fx, uint64_t* words, uint64_t max_words, uint64_t* offsets,
                                   uint8_t* flags, uint64_t max_reads, uint64_t* n_real);
/* n_parsed(), n_skipped(), is_complete() (readers.hh:205-215). */
int gt_fastx_stats(const gt_fastx* fx, uint64_t* n_parsed, uint64_t* n_skipped, int* is_complete);
/* FileProcessor<InserterProcessor<dBG>>::process / advance (processors.hh:112-127, 208-229, 304-331):
 * streams up to max_reads records (0 = the rest of the file) from the parser into the graph --
 * batch n+1 is parsed into pinned memory while batch n is copied, packed, hashed and inserted.
 * Returns the k-mers consumed (the processor's "time"); *n_seqs = records processed (records shorter
 * than K count and contribute 0 k-mers, as InserterProcessor::process_sequence swallows the
 * SequenceLengthException). */
int64_t gt_insert_fastx(gt_storage* st, int shifter, int K, gt_fastx* fx, int mode, uint64_t max_reads,
                        uint64_t* n_seqs);

/* FileProcessor::advance (processors.hh:208-229) in the reference's own "time" unit: consume records until the k-mers
 * processed by this call reach interval_kmers (IntervalCounter::poll, metrics.hh:129-138: the record that crosses the
 * threshold is the last one consumed) or the file ends.  Returns the k-mers consumed; *n_seqs = records consumed;
 * *remaining = 1 when the interval ended the call, 0 at the end of the file. */
int64_t gt_insert_fastx_advance(gt_storage* st, int shifter, int K, gt_fastx* fx, int mode, uint64_t interval_kmers,
                                uint64_t* n_seqs, int* remaining);

/* ---- SourmashSketch (sketches/sourmash_sketch.hh:24-82) -------------------------------- */
/* Sketch(n, K, is_protein=false, dayhoff=false, hp=false, seed, scaled):
 * max_hash_from_scaled (:52-61) is applied by the caller via gt_max_hash_from_scaled. */
uint64_t gt_max_hash_from_scaled(uint64_t scaled);
gt_sketch* gt_sketch_create(uint32_t num, int K, uint32_t seed, uint64_t max_hash);
void gt_sketch_destroy(gt_sketch* sk);
/* Sketch::insert_sequence over a batch (:71-81; add_sequence(seq, force=true)): windows with
 * a byte outside ACGTacgt are skipped.  Returns sum of len-K+1 over reads with len >= K. */
int64_t gt_sketch_add_sequences(gt_sketch* sk, const char* bases, const uint64_t* offsets,
                                uint64_t n_reads);
/* Same for reads already resident in HBM (ASCII d_bases, uint64 d_offsets starting at 0): pack + sketch on the
 * library's compute stream; one 8-byte read-back. */
int64_t gt_sketch_add_sequences_dev(gt_sketch* sk, const void* d_bases, const void* d_offsets, uint64_t n_reads,
                                    uint64_t n_bases);
/* MinHash::add_hash / merge (sourmash.hpp:80, :97) */
int gt_sketch_add_hashes(gt_sketch* sk, const uint64_t* hashes, uint64_t n);
/* Device-resident forms for the union of a sharded sketch (every rank sketches its reads; the sets are united by
 * all-gather): the number of hashes the set holds now, the hashes (unsorted) into a device buffer, and hashes added
 * from a device buffer. */
int64_t gt_sketch_live_count(gt_sketch* sk);
int64_t gt_sketch_export_dev(gt_sketch* sk, uint64_t* d_out, uint64_t capacity);
int gt_sketch_add_hashes_dev(gt_sketch* sk, const uint64_t* d_hashes, uint64_t n);
int gt_sketch_merge(gt_sketch* dst, gt_sketch* src);
/* MinHash::size / mins (sourmash.hpp:116, :151-157): ascending, duplicate-free. */
int64_t gt_sketch_size(gt_sketch* sk);
int64_t gt_sketch_mins(gt_sketch* sk, uint64_t* out, uint64_t capacity);
/* MinHash::count_common (sourmash.hpp:102-106): |A n B| of two sketches with equal parameters. */
int64_t gt_sketch_count_common(gt_sketch* a, gt_sketch* b);
/* Empty the sketch (parameters kept). */
int gt_sketch_reset(gt_sketch* sk);

#ifdef __cplusplus
}
#endif
#endif /* GOETIA_B200_H */
