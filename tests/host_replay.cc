// Host-only harness for the native replay code (zk-lisp_b200/csrc/api_replay.inl): the proof parser and the merge-plan builder are
// plain C++ and are compiled here with g++ so that the CPU suite can compare them with the Python restatement (replay.py / step.py).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#define ZKL_REPLAY_HOST_ONLY
#include "../zk-lisp_b200/csrc/api_replay.inl"

extern "C" {
// fields[0..14): W, aux, logn, num_constraints, q, blowup, grind, ext, folding, rem_deg, parts, rate, unique, num_layers;
// spans[2k], spans[2k+1] = offset, length of: commitments, tq_values, tq_paths, cq_values, cq_paths, ood_trace, ood_cons, remainder,
// then (values, paths) per FRI layer.
int hr_parse(const uint8_t *proof, size_t len, uint64_t *fields, uint64_t *nonce, uint64_t *spans, size_t max_spans) {
  rp_view v;
  if (!rp_parse(proof, len, v)) return 1;
  uint64_t f[14] = {v.W, v.aux, v.logn, v.num_constraints, v.q, v.blowup, v.grind, v.ext, v.folding, v.rem_deg, v.parts, v.rate, v.unique, v.layers.size()};
  memcpy(fields, f, sizeof f);
  *nonce = v.nonce;
  std::vector<rp_span> s = {v.commitments, v.tq_values, v.tq_paths, v.cq_values, v.cq_paths, v.ood_trace, v.ood_cons, v.remainder};
  for (auto &l : v.layers) { s.push_back(l.first); s.push_back(l.second); }
  if (s.size() > max_spans) return 2;
  for (size_t k = 0; k < s.size(); k++) { spans[2 * k] = (uint64_t)(s[k].p - proof); spans[2 * k + 1] = s[k].n; }
  return 0;
}
// Plan of `num` openings given as serialized batch proofs; leaves are 32-byte digests, positions per opening.
// Outputs the flattened schedule like MergePlan.flatten(): returns the number of levels, or -1 - (index of the malformed opening).
int hr_plan(size_t num, const uint8_t *const *paths, const size_t *path_len, const uint8_t *const *leaves, const uint32_t *const *positions,
            const size_t *num_positions, uint8_t *known, size_t *num_known, uint32_t *left, uint32_t *right, uint32_t *counts, uint32_t *want) {
  rp_plan plan;
  static const uint8_t zero[32] = {0};
  for (size_t o = 0; o < num; o++) {
    int depth; std::vector<std::vector<const uint8_t *>> lists;
    std::vector<const uint8_t *> lv; std::vector<uint32_t> pos(positions[o], positions[o] + num_positions[o]);
    for (size_t k = 0; k < num_positions[o]; k++) lv.push_back(leaves[o] + 32 * k);
    rp_span s; s.p = paths[o]; s.n = path_len[o];
    if (!rp_read_batch_proof(s, depth, lists) || !plan.add_opening(depth, lists, lv, pos, zero, o, 0)) return -1 - (int)o;
  }
  *num_known = plan.known.size();
  for (size_t k = 0; k < plan.known.size(); k++) memcpy(known + 32 * k, plan.known[k], 32);
  std::vector<uint32_t> base(plan.levels.size());
  uint32_t acc = (uint32_t)plan.known.size();
  for (size_t l = 0; l < plan.levels.size(); l++) { base[l] = acc; acc += (uint32_t)plan.levels[l].size(); counts[l] = (uint32_t)plan.levels[l].size(); }
  auto index = [&](rp_node x) { return x.level < 0 ? x.idx : base[(size_t)x.level] + x.idx; };
  size_t at = 0;
  for (auto &lv : plan.levels) for (auto &m : lv) { left[at] = index(m.first); right[at] = index(m.second); at++; }
  for (size_t k = 0; k < plan.roots.size(); k++) want[k] = index(plan.roots[k].node);
  return (int)plan.levels.size();
}
}
