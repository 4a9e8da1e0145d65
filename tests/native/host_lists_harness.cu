// Host-only harness of the list builders in chalte.cu (no GPU needed: nothing here touches the CUDA runtime).
// It includes the engine's translation unit, fills the host-side fields the builders read from two binary files
// (frequency-sorted selected lines, ascending channel frequencies: float64) and checks, for each half-width given on
// the command line, the invariants the device kernels rely on:
//   groups / records / tiles (make_group_lists) and the span -> segment table of the one-pass channel-stream kernel
//   (make_span_table): every active channel is produced by exactly one (span, segment, group, lane), no inactive
//   channel by any; a group's records and a record's line are found where the kernel looks for them.
// usage: host_lists_harness lines.bin freq.bin mask_centre hv [hv ...]   -> one JSON line per hv, exit code 0 / 1
//        (lines.bin.mol, if present: molecule id per line as float64 -- a joint fit of several molecules)
#include "../../cha1_mcmc_b200/csrc/chalte.cu"

#include <fstream>

static std::vector<double> read_f64(const char* f) {
  std::ifstream s(f, std::ios::binary | std::ios::ate);
  if (!s) return {};
  const size_t n = (size_t)s.tellg();
  s.seekg(0);
  std::vector<double> v(n / 8);
  s.read((char*)v.data(), (std::streamsize)n);
  return v;
}

#define REQUIRE(cond, msg)                                                              \
  do {                                                                                  \
    if (!(cond)) { fprintf(stderr, "FAILED (hv %.4f): %s [%s]\n", hv, msg, #cond); return 1; } \
  } while (0)

static int check(cha_engine* h, double hv) {
  HostLists& L = h->host_wide;
  if (make_group_lists(h, hv, L)) { fprintf(stderr, "make_group_lists: %s\n", h->err.c_str()); return 1; }
  const size_t C = h->xs.size(), A = L.act_ch.size(), G = L.gblk.size(), T = L.tiles.size();
  REQUIRE(L.recs.size() >= 1, "trailing look-ahead record");
  const size_t R = L.recs.size() - 1;
  // ---- groups: consecutive active channels, <= 8 each, records contiguous, every record's window covers the group
  size_t a = 0;
  for (size_t g = 0; g < G; ++g) {
    const GroupBlk& gb = L.gblk[g];
    int n = 0;
    for (int j = 0; j < kGroupCh; ++j) if (gb.opos[j] >= 0) { REQUIRE(j == n, "live lanes first"); ++n; }
    REQUIRE(n >= 1, "group without a channel");
    for (int j = 0; j < n; ++j) { REQUIRE(a < A && gb.opos[j] == L.act_ch[a], "groups walk the active channels in order"); ++a; }
    REQUIRE(L.grp_c0[g] == gb.opos[0] && L.grp_c1[g] == gb.opos[n - 1], "group channel range");
    int nrec = 0;
    for (int m = 0; m < kMaxM; ++m) nrec += gb.nrec[m];
    REQUIRE(L.grp_rec1[g] - L.grp_rec0[g] == nrec, "record count of the group");
    REQUIRE(g + 1 == G || L.grp_rec1[g] == L.grp_rec0[g + 1], "records of consecutive groups are contiguous");
    for (int q = L.grp_rec0[g]; q < L.grp_rec1[g]; ++q) {
      const int i = L.recs[q].line;
      REQUIRE(i >= 0 && i < (int)h->l_nu.size(), "record line id");
      REQUIRE(L.wa[i] <= gb.opos[n - 1] && L.wb[i] > gb.opos[0], "the record's line window meets the group");
    }
    // molecule after molecule, nrec[m] records each (what the multi-molecule fast path's flat record loop assumes)
    {
      int q = L.grp_rec0[g];
      for (int m = 0; m < kMaxM; ++m)
        for (int k = 0; k < gb.nrec[m]; ++k, ++q) {
          REQUIRE(h->l_mol[L.recs[q].line] == m, "records of a group are ordered by molecule");
          REQUIRE(k == 0 || L.recs[q].line > L.recs[q - 1].line, "and by line within a molecule");
        }
    }
  }
  REQUIRE(a == A, "every active channel is in a group");
  REQUIRE(G == 0 || (size_t)L.grp_rec1[G - 1] == R, "records end with the last group");
  // ---- tiles: consecutive groups, limits, relative offsets
  size_t g_next = 0;
  for (size_t t = 0; t < T; ++t) {
    const TileG& tl = L.tiles[t];
    REQUIRE((size_t)tl.g0 == g_next && tl.ng >= 1 && tl.ng <= kTileMaxGroups, "tiles partition the groups");
    g_next += (size_t)tl.ng;
    REQUIRE(tl.rec_begin == L.grp_rec0[tl.g0], "tile record start");
    for (int g = tl.g0; g < tl.g0 + tl.ng; ++g) {
      REQUIRE(L.gblk[g].rec_off == L.grp_rec0[g] - tl.rec_begin, "group.rec_off is relative to the tile");
      for (int q = L.grp_rec0[g]; q < L.grp_rec1[g]; ++q) {
        REQUIRE(L.recs[q].lloc == (L.recs[q].line - tl.line0) * kWalkersPerBlock, "record.lloc is relative to the tile");
        REQUIRE(L.recs[q].line >= tl.line0 && L.recs[q].line < tl.line0 + tl.nline, "tile line range");
      }
    }
    REQUIRE(tl.inv_hs == 1.0 / tl.hs, "1/hs");
  }
  REQUIRE(g_next == G, "every group is in a tile");
  // ---- span table (the engine builds it only when every tile can be staged: build_pairs)
  if (L.n_unstaged > 0) {
    printf("{\"hv\": %.4f, \"channels\": %zu, \"active\": %zu, \"groups\": %zu, \"records\": %zu, \"tiles\": %zu, \"pairs\": %lld, "
           "\"unstaged_tiles\": %lld}\n", hv, C, A, G, R, T, (long long)L.P, (long long)L.n_unstaged);
    return 0;
  }
  std::vector<int> soff; std::vector<SpanSeg> segs;
  bool sparse = false;
  const bool fits = make_span_table(L, C, soff, segs, sparse);
  const size_t ns = (C + kSpanCh - 1) / kSpanCh;
  REQUIRE(soff.size() == ns + 1 && soff[0] == 0 && (size_t)soff[ns] + 1 == segs.size(), "offsets");
  std::vector<unsigned char> cover(C, 0);
  size_t multi = 0, nonempty = 0;
  for (size_t s = 0; s < ns; ++s) {
    REQUIRE(soff[s + 1] >= soff[s], "offsets ascend");
    nonempty += soff[s + 1] > soff[s];
    multi += soff[s + 1] - soff[s] > 1;
    const int c0 = (int)(s * kSpanCh), c1 = (int)std::min<size_t>(C, (s + 1) * (size_t)kSpanCh);
    int prev_g_end = -1;
    for (int k = soff[s]; k < soff[s + 1]; ++k) {
      const SpanSeg& sg = segs[k];
      REQUIRE(sg.tile >= 0 && (size_t)sg.tile < T, "segment tile");
      const TileG& tl = L.tiles[sg.tile];
      REQUIRE(sg.g_lo >= tl.g0 && sg.g_lo + sg.g_n <= tl.g0 + tl.ng && sg.g_n >= 1 && sg.g_n <= kSegGroups, "segment groups lie in its tile");
      REQUIRE(sg.g_lo >= prev_g_end, "segments of a span ascend and do not overlap");
      prev_g_end = sg.g_lo + sg.g_n;
      REQUIRE(sg.r_lo == L.grp_rec0[sg.g_lo] && sg.r_n == L.grp_rec1[sg.g_lo + sg.g_n - 1] - sg.r_lo, "segment records");
      if (fits) REQUIRE(sg.r_n <= kSegRecs, "segment fits the record staging area");
      REQUIRE(sg.l_n <= kTileMaxLines, "segment fits the strength staging area");
      REQUIRE(sg.inv_hs == (float)(1.0 / tl.hs), "segment 1/hs");
      for (int g = sg.g_lo; g < sg.g_lo + sg.g_n; ++g) {
        const GroupBlk& gb = L.gblk[g];
        int nrec = 0;
        for (int m = 0; m < kMaxM; ++m) nrec += gb.nrec[m];
        const int r = gb.rec_off + sg.rec_shift;                       // where the kernel looks for the group's records
        REQUIRE(r >= 0 && r + nrec <= sg.r_n && sg.r_lo + r == L.grp_rec0[g], "staged record index of the group");
        for (int q = L.grp_rec0[g]; q < L.grp_rec1[g]; ++q) {
          const int row = L.recs[q].lloc / kWalkersPerBlock + sg.line_shift;       // staged strength row of the record
          REQUIRE(row >= 0 && row < sg.l_n && sg.l_lo + row == L.recs[q].line, "staged strength row of the record");
        }
        for (int j = 0; j < kGroupCh; ++j) {
          const int o = gb.opos[j];
          if (o >= c0 && o < c1) { REQUIRE(cover[o] < 255, "cover"); cover[o]++; }
        }
      }
    }
  }
  size_t ai = 0;
  for (size_t j = 0; j < C; ++j) {
    const bool active = ai < A && (size_t)L.act_ch[ai] == j;
    if (active) ++ai;
    REQUIRE(cover[j] == (active ? 1 : 0), "every active channel is produced exactly once, no inactive one");
  }
  printf("{\"hv\": %.4f, \"channels\": %zu, \"active\": %zu, \"groups\": %zu, \"records\": %zu, \"tiles\": %zu, \"pairs\": %lld, "
         "\"spans\": %zu, \"segments\": %zu, \"nonempty_spans\": %zu, \"multi_segment_spans\": %zu, \"fits\": %s, \"sparse\": %s}\n",
         hv, C, A, G, R, T, (long long)L.P, ns, segs.size() - 1, nonempty, multi, fits ? "true" : "false", sparse ? "true" : "false");
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s lines.bin freq.bin mask_centre hv [hv ...]\n", argv[0]); return 2; }
  cha_engine* h = new cha_engine();
  h->md.M = 1; h->md.K = 1; h->md.mc = atof(argv[3]); h->md.eps = 1e-10; h->md.dish = 100.0;
  h->l_nu = read_f64(argv[1]); h->l_mol.assign(h->l_nu.size(), 0);
  {
    // optional: molecule id per line (float64 file <lines>.mol) for joint fits
    const std::vector<double> mol = read_f64((std::string(argv[1]) + ".mol").c_str());
    if (mol.size() == h->l_nu.size()) {
      int M = 1;
      for (size_t i = 0; i < mol.size(); ++i) { h->l_mol[i] = (int)mol[i]; M = std::max(M, h->l_mol[i] + 1); }
      h->md.M = M;
    }
  }
  h->xs = read_f64(argv[2]);
  const size_t C = h->xs.size();
  h->ys.assign(C, 0.01); h->ws.assign(C, 1e4); h->iss.assign(C, 100.0);
  h->perm.resize(C);
  for (size_t j = 0; j < C; ++j) h->perm[j] = (int)j;
  h->y2w_prefix.assign(C + 1, 0.0);
  for (size_t j = 0; j < C; ++j) h->y2w_prefix[j + 1] = h->y2w_prefix[j] + h->ys[j] * h->ys[j] * h->ws[j];
  int rc = 0;
  for (int k = 4; k < argc; ++k) rc |= check(h, atof(argv[k]));
  return rc;
}
