/*
 * oracle/oracle.c — plain-C restatement of the reference's shared_tree path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  It restates WHAT the
 * reference computes as one global level-by-level pass with integer
 * arithmetic; it shares no code with the reference and none with the CUDA
 * product.  Every function cites the reference lines it follows
 * (paths relative to the reference root).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* leaves                                                                   */
/* ------------------------------------------------------------------------ */

/* src/dna.cpp:25-49 (to_nac) with the code values of include/dna.h:20-32. */
int orc_code(int ch) {
  if (ch >= 'a' && ch <= 'z') ch -= 32; /* std::toupper in the "C" locale */
  switch (ch) {
    case 'A': return 0x1;
    case 'C': return 0x2;
    case 'G': return 0x4;
    case 'T': return 0x8;
    case 'R': return 0x3;
    case 'Y': return 0xC;
    case 'K': return 0x7;
    case 'M': return 0xE;
    case 'S': return 0x0;
    case 'W': return 0x9;
    case 'B': return 0x5;
    case 'D': return 0xB;
    case 'H': return 0xD;
    case 'V': return 0xA;
    case 'N': return 0x6;
    case '-': return 0xF;
    default: return -1;
  }
}

static const char k_letters[16] = {'S', 'A', 'C', 'R', 'G', 'B', 'N', 'K',
                                   'T', 'W', 'V', 'D', 'Y', 'H', 'M', '-'};

/* src/dna.cpp:79-84, :187-197: nucleotide i occupies bits 4i..4i+3. */
uint64_t orc_pack(const char* text, int S, int* bad) {
  uint64_t v = 0;
  if (bad) *bad = -1;
  for (int i = 0; i < S; ++i) {
    const int c = orc_code((unsigned char)text[i]);
    if (c < 0) {
      int up = (unsigned char)text[i];
      if (up >= 'a' && up <= 'z') up -= 32;
      if (bad && *bad < 0) *bad = up;
      continue;
    }
    v |= (uint64_t)c << (4 * i);
  }
  return v;
}

void orc_leaf_to_ascii(uint64_t v, int S, char* out) {
  for (int i = 0; i < S; ++i) out[i] = k_letters[(v >> (4 * i)) & 0xF];
}

/* src/dna.cpp:104-111: reverse the four bits of every nibble. */
uint64_t orc_transposed(uint64_t v) {
  v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
  v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
  return v;
}

/* src/dna.cpp:116-121: nibble i <- nibble S-1-i. */
uint64_t orc_mirrored(uint64_t v, int S) {
  uint64_t r = 0;
  for (int i = 0; i < S; ++i) r |= ((v >> (4 * (S - 1 - i))) & 0xFull) << (4 * i);
  return r;
}

/* include/dna.h:51: transposed().mirrored() */
uint64_t orc_inverted(uint64_t v, int S) { return orc_mirrored(orc_transposed(v), S); }

/* src/dna.cpp:135-143: minimum of (value, mirror, transpose) over the four
 * similarity transforms; invariant = (x == mirrored(x)) (include/dna.h:52). */
uint64_t orc_leaf_canonical(uint64_t v, int S, int* flags) {
  const uint64_t cand[4] = {v, orc_transposed(v), orc_mirrored(v, S), orc_inverted(v, S)};
  const int cflag[4] = {0, 2, 1, 3}; /* bit0 mirror, bit1 transpose */
  /* tuple order is (value, mirror, transpose): candidates listed in
   * increasing (mirror, transpose) order, strict '<' keeps the earliest. */
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (cand[i] < cand[best]) best = i;
  const int inv = (v == cand[2]);
  *flags = cflag[best] | (inv << 2);
  return cand[best];
}

/* ------------------------------------------------------------------------ */
/* pointers and nodes                                                       */
/* ------------------------------------------------------------------------ */

/* src/shared_tree.cpp:86-87: mirror is dropped on invariant targets. */
uint32_t orc_ptr(uint64_t index, int m, int t, int inv) {
  uint32_t raw = (uint32_t)index & ORC_IDX_MASK;
  if (m && !inv) raw |= ORC_MIRROR;
  if (t) raw |= ORC_TRANSPOSE;
  if (inv) raw |= ORC_INVARIANT;
  return raw;
}

/* src/shared_tree.cpp:76-80: pointer{other, mirror, transpose}. `other !=
 * nullptr` compares the 31-bit keys (include/shared_tree.h:54-57). */
uint32_t orc_compose(uint32_t p, int m, int t) {
  const int pm = (p >> 29) & 1, pt = (p >> 30) & 1, pinv = (p >> 31) & 1;
  const int is_null = (p & ORC_KEY31) == ORC_IDX_MASK;
  uint32_t raw = p & ORC_IDX_MASK;
  if ((m != pm) && !pinv) raw |= ORC_MIRROR;
  if ((t != pt) && !is_null) raw |= ORC_TRANSPOSE;
  if (pinv) raw |= ORC_INVARIANT;
  return raw;
}

static uint64_t key64(uint32_t l, uint32_t r) {
  return ((uint64_t)(l & ORC_KEY31) << 32) | (uint64_t)(r & ORC_KEY31);
}

/* include/shared_tree.h:111-126 (variants, canonical) and
 * src/shared_tree.cpp:662-672 (invariant = left == right.mirrored()).
 * Returns bit0 mirror, bit1 transpose, bit2 invariant. */
int orc_node_canonical(uint32_t l, uint32_t r, uint32_t out[2]) {
  const uint32_t cl[4] = {l, orc_compose(l, 0, 1), orc_compose(r, 1, 0), orc_compose(r, 1, 1)};
  const uint32_t cr[4] = {r, orc_compose(r, 0, 1), orc_compose(l, 1, 0), orc_compose(l, 1, 1)};
  const int cflag[4] = {0, 2, 1, 3}; /* (m,t) = (0,0) (0,1) (1,0) (1,1) */
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (key64(cl[i], cr[i]) < key64(cl[best], cr[best])) best = i;
  out[0] = cl[best];
  out[1] = cr[best];
  const int inv = ((l & ORC_KEY31) == (orc_compose(r, 1, 0) & ORC_KEY31));
  return cflag[best] | (inv << 2);
}

/* src/shared_tree.cpp:25-58: segment starts 0 / 16 / 4112 / 1052688. */
static void compress_index(uint32_t index, int* segment, uint32_t* offset) {
  if (index == ORC_IDX_MASK) {
    *segment = 3;
    *offset = 0xfffffffu;
  } else if (index < 16u) {
    *segment = 0;
    *offset = index;
  } else if (index < 4112u) {
    *segment = 1;
    *offset = index - 16u;
  } else if (index < 1052688u) {
    *segment = 2;
    *offset = index - 4112u;
  } else {
    *segment = 3;
    *offset = index - 1052688u;
  }
}

/* src/shared_tree.cpp:122-125 */
int orc_ptr_bytes(uint32_t raw) {
  int segment;
  uint32_t offset;
  compress_index(raw & ORC_IDX_MASK, &segment, &offset);
  return segment + 1;
}

/* src/shared_tree.cpp:133-142 */
int orc_ptr_serialize(uint32_t raw, uint8_t* out) {
  static const int bits[4] = {4, 12, 20, 28};
  int segment;
  uint32_t offset;
  compress_index(raw & ORC_IDX_MASK, &segment, &offset);
  int shift = bits[segment] - 4;
  int n = 0;
  out[n++] = (uint8_t)((offset >> shift) | (((raw >> 29) & 1) << 4) | (((raw >> 30) & 1) << 5) |
                       ((uint32_t)segment << 6));
  for (shift -= 8; shift >= 0; shift -= 8) out[n++] = (uint8_t)(offset >> shift);
  return n;
}

/* src/shared_tree.cpp:147-163; returns bytes consumed, 0 if truncated. */
static int ptr_deserialize(const uint8_t* in, uint64_t avail, uint32_t* raw) {
  static const int bits[4] = {4, 12, 20, 28};
  static const uint32_t start[4] = {0u, 16u, 4112u, 1052688u};
  if (avail < 1) return 0;
  const int segment = (in[0] >> 6) & 3;
  if (avail < (uint64_t)segment + 1) return 0;
  const int t = (in[0] >> 5) & 1, m = (in[0] >> 4) & 1;
  int shift = bits[segment] - 4;
  uint32_t offset = ((uint32_t)in[0] & 0xFu) << shift;
  int n = 1;
  for (shift -= 8; shift >= 0; shift -= 8) offset |= (uint32_t)in[n++] << shift;
  const uint32_t index = (segment == 3 && offset == 0xfffffffu) ? ORC_IDX_MASK : start[segment] + offset;
  *raw = orc_ptr(index, m, t, 0); /* invariant is never stored (:162) */
  return n;
}

/* ------------------------------------------------------------------------ */
/* ingest                                                                   */
/* ------------------------------------------------------------------------ */

/* src/fasta_reader.cpp:40-68.  Per getline round: if the next byte is '>' or
 * '\n' one line is discarded (:47-49), then one line is read as data (:50)
 * whatever it starts with.  Restated as a two-state machine over lines:
 * CHECK -> (line starts '>' or is blank: skip it, next line UNCHECKED);
 * UNCHECKED -> the line is data. Newlines are never data. */
uint64_t orc_fasta_body(const char* text, uint64_t len, char* out) {
  uint64_t pos = 0, n = 0;
  int unchecked = 0;
  while (pos < len) {
    uint64_t end = pos;
    while (end < len && text[end] != '\n') ++end;
    const int blank = (end == pos) && (end < len); /* an immediate '\n' */
    if (!unchecked && (text[pos] == '>' || blank)) {
      unchecked = 1;
    } else {
      memcpy(out + n, text + pos, end - pos);
      n += end - pos;
      unchecked = 0;
    }
    pos = end + 1;
  }
  return n;
}

/* src/fasta_reader.cpp:59-67: drop the last len mod S bytes, then pack. */
uint64_t orc_fasta_to_leaves(const char* text, uint64_t len, int S, uint64_t* out, uint64_t cap,
                             int* bad) {
  char* body = (char*)malloc(len ? len : 1);
  const uint64_t n = orc_fasta_body(text, len, body);
  const uint64_t leaves = n / (uint64_t)S;
  if (bad) *bad = -1;
  for (uint64_t i = 0; i < leaves; ++i) {
    int b;
    const uint64_t v = orc_pack(body + i * (uint64_t)S, S, &b);
    if (b >= 0) {
      if (bad) *bad = b;
      free(body);
      return UINT64_MAX;
    }
    if (i < cap) out[i] = v;
  }
  free(body);
  return leaves;
}

/* ------------------------------------------------------------------------ */
/* first-occurrence dictionary (stands in for phmap::parallel_flat_hash_map; */
/* results never depend on it: src/shared_tree.cpp:632-633, :666-668)        */
/* ------------------------------------------------------------------------ */

typedef struct dict {
  uint64_t cap; /* power of two */
  uint64_t used;
  uint64_t* keys;
  uint32_t* ids; /* id + 1, 0 = empty */
} dict;

static void dict_init(dict* d, uint64_t expect) {
  d->cap = 1024;
  while (d->cap < expect * 2) d->cap <<= 1;
  d->used = 0;
  d->keys = (uint64_t*)malloc(d->cap * sizeof(uint64_t));
  d->ids = (uint32_t*)calloc(d->cap, sizeof(uint32_t));
}

static void dict_free(dict* d) {
  free(d->keys);
  free(d->ids);
}

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

/* Returns the id of key, inserting it with id = d->used if absent. */
static uint32_t dict_emplace(dict* d, uint64_t key, int* inserted) {
  uint64_t s = mix64(key) & (d->cap - 1);
  for (;;) {
    if (d->ids[s] == 0) {
      d->keys[s] = key;
      d->ids[s] = (uint32_t)(d->used + 1);
      *inserted = 1;
      return (uint32_t)d->used++;
    }
    if (d->keys[s] == key) {
      *inserted = 0;
      return d->ids[s] - 1;
    }
    s = (s + 1) & (d->cap - 1);
  }
}

/* ------------------------------------------------------------------------ */
/* build                                                                    */
/* ------------------------------------------------------------------------ */

/* One global level-by-level pass.  Equivalent to the reference's segmented
 * streaming (include/shared_tree.h:305-316, src/shared_tree.cpp:677-763)
 * because its segments are powers of two. */
orc_tree* orc_build_levels(const uint64_t* leaves, uint64_t n, int S, uint32_t* out, uint64_t cap,
                           uint64_t* level_sizes, int max_levels, int* n_levels) {
  if (n == 0) return NULL;
  orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
  t->S = S;
  uint64_t written = 0;
  int levels = 0;

  /* leaf level: src/shared_tree.cpp:630-637 (emplace_leaf) */
  uint32_t* cur = (uint32_t*)malloc(n * sizeof(uint32_t));
  t->leaves = (uint64_t*)malloc(n * sizeof(uint64_t));
  {
    dict d;
    dict_init(&d, n);
    for (uint64_t i = 0; i < n; ++i) {
      int flags, inserted;
      const uint64_t canon = orc_leaf_canonical(leaves[i], S, &flags);
      /* key 0 is a legal leaf (all 'S'); the dict stores id+1 so no sentinel */
      const uint32_t id = dict_emplace(&d, canon, &inserted);
      if (inserted) t->leaves[id] = canon;
      cur[i] = orc_ptr(id, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1);
    }
    t->n_leaves = d.used;
    dict_free(&d);
  }
  t->leaves = (uint64_t*)realloc(t->leaves, (t->n_leaves ? t->n_leaves : 1) * sizeof(uint64_t));

  /* node levels: include/shared_tree.h:282-299, src/shared_tree.cpp:697-712,
   * emplace_node :662-672.  At least one node layer exists (a single leaf is
   * wrapped as node{leaf, null}). */
  int cap_layers = 64;
  t->layer_size = (uint64_t*)calloc(cap_layers, sizeof(uint64_t));
  t->layers = (uint32_t**)calloc(cap_layers, sizeof(uint32_t*));
  uint64_t cur_n = n;
  do {
    const uint64_t next_n = (cur_n + 1) / 2;
    uint32_t* next = (uint32_t*)malloc(next_n * sizeof(uint32_t));
    uint32_t* nodes = (uint32_t*)malloc(next_n * 2 * sizeof(uint32_t));
    dict d;
    dict_init(&d, next_n);
    for (uint64_t i = 0; i < next_n; ++i) {
      const uint32_t l = cur[2 * i];
      const uint32_t r = (2 * i + 1 < cur_n) ? cur[2 * i + 1] : ORC_NULL;
      uint32_t c[2];
      const int flags = orc_node_canonical(l, r, c);
      int inserted;
      const uint32_t id = dict_emplace(&d, key64(c[0], c[1]), &inserted);
      if (inserted) {
        nodes[2 * id] = c[0];
        nodes[2 * id + 1] = c[1];
      }
      next[i] = orc_ptr(id, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1);
    }
    t->layer_size[t->n_layers] = d.used;
    t->layers[t->n_layers] = (uint32_t*)realloc(nodes, d.used * 2 * sizeof(uint32_t));
    t->n_layers++;
    dict_free(&d);
    if (out) {
      if (levels < max_levels) level_sizes[levels] = next_n;
      for (uint64_t i = 0; i < next_n; ++i)
        if (written + i < cap) out[written + i] = next[i];
      written += next_n;
    }
    ++levels;
    free(cur);
    cur = next;
    cur_n = next_n;
  } while (cur_n > 1);
  t->root = cur[0];
  free(cur);
  if (n_levels) *n_levels = levels;
  return t;
}

orc_tree* orc_build(const uint64_t* leaves, uint64_t n, int S) {
  return orc_build_levels(leaves, n, S, NULL, 0, NULL, 0, NULL);
}

void orc_free(orc_tree* t) {
  if (!t) return;
  for (int k = 0; k < t->n_layers; ++k) free(t->layers[k]);
  free(t->layers);
  free(t->layer_size);
  free(t->leaves);
  free(t);
}

uint64_t orc_tree_leaf_count(const orc_tree* t) { return t->n_leaves; }
int orc_tree_layers(const orc_tree* t) { return t->n_layers; }
uint64_t orc_tree_layer_size(const orc_tree* t, int layer) { return t->layer_size[layer]; }
const uint64_t* orc_tree_leaves(const orc_tree* t) { return t->leaves; }
const uint32_t* orc_tree_layer(const orc_tree* t, int layer) { return t->layers[layer]; }
uint32_t orc_tree_root(const orc_tree* t) { return t->root; }

uint64_t orc_node_count(const orc_tree* t) {
  uint64_t s = 0;
  for (int k = 0; k < t->n_layers; ++k) s += t->layer_size[k];
  return s;
}

/* ------------------------------------------------------------------------ */
/* histogram + frequency sort                                               */
/* ------------------------------------------------------------------------ */

/* src/shared_tree.cpp:316-326; `layer` is the PARENT node layer, the result
 * counts references into its child layer (leaves for layer 0). */
void orc_histogram(const orc_tree* t, int layer, uint64_t* out) {
  const uint64_t child_n = layer == 0 ? t->n_leaves : t->layer_size[layer - 1];
  memset(out, 0, child_n * sizeof(uint64_t));
  for (uint64_t i = 0; i < 2 * t->layer_size[layer]; ++i) {
    const uint32_t p = t->layers[layer][i];
    if ((p & ORC_KEY31) != ORC_IDX_MASK) out[p & ORC_IDX_MASK]++;
  }
}

/* stable merge sort of indices by descending frequency
 * (std::stable_sort, src/shared_tree.cpp:413, :430) */
static void stable_sort_desc(uint64_t* idx, uint64_t n, const uint64_t* freq) {
  uint64_t* tmp = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  for (uint64_t w = 1; w < n; w *= 2) {
    for (uint64_t lo = 0; lo < n; lo += 2 * w) {
      const uint64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      uint64_t a = lo, b = mid, o = lo;
      while (a < mid && b < hi) tmp[o++] = (freq[idx[b]] > freq[idx[a]]) ? idx[b++] : idx[a++];
      while (a < mid) tmp[o++] = idx[a++];
      while (b < hi) tmp[o++] = idx[b++];
    }
    memcpy(idx, tmp, n * sizeof(uint64_t));
  }
  free(tmp);
}

/* src/shared_tree.cpp:409-437: sort child layer `child` (0 = leaves,
 * c>0 = node layer c-1) by reference count from parent node layer `child`;
 * permute it (:371) and rewire the parents' indices, flags kept (:383-403). */
static void sort_child_layer(orc_tree* t, int child) {
  const int parent = child;
  const uint64_t n = child == 0 ? t->n_leaves : t->layer_size[child - 1];
  uint64_t* freq = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  uint64_t* order = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  uint64_t* newpos = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  orc_histogram(t, parent, freq);
  for (uint64_t i = 0; i < n; ++i) order[i] = i;
  stable_sort_desc(order, n, freq);
  for (uint64_t i = 0; i < n; ++i) newpos[order[i]] = i; /* invert_indices :360 */
  if (child == 0) {
    uint64_t* moved = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) moved[newpos[i]] = t->leaves[i];
    free(t->leaves);
    t->leaves = moved;
  } else {
    uint32_t* moved = (uint32_t*)malloc((n ? n : 1) * 2 * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; ++i) {
      moved[2 * newpos[i]] = t->layers[child - 1][2 * i];
      moved[2 * newpos[i] + 1] = t->layers[child - 1][2 * i + 1];
    }
    free(t->layers[child - 1]);
    t->layers[child - 1] = moved;
  }
  for (uint64_t i = 0; i < 2 * t->layer_size[parent]; ++i) {
    const uint32_t p = t->layers[parent][i];
    if ((p & ORC_KEY31) == ORC_IDX_MASK) continue; /* old.empty() */
    t->layers[parent][i] = (p & ~ORC_IDX_MASK) | (uint32_t)newpos[p & ORC_IDX_MASK];
  }
  free(freq);
  free(order);
  free(newpos);
}

/* src/shared_tree.cpp:443-483: leaves and node layers 0..L-2 are each sorted
 * once; the top layer and the root are untouched. Order between layers does
 * not matter (each histogram is invariant under permuting its parent). */
void orc_sort(orc_tree* t) {
  for (int child = 0; child < t->n_layers; ++child) sort_child_layer(t, child);
}

/* ------------------------------------------------------------------------ */
/* serialization                                                            */
/* ------------------------------------------------------------------------ */

/* src/shared_tree.cpp:488-496 */
uint64_t orc_bytes(const orc_tree* t) {
  uint64_t total = (uint64_t)orc_ptr_bytes(t->root) + 8 + t->n_leaves * (uint64_t)((t->S + 1) / 2);
  for (int k = 0; k < t->n_layers; ++k) {
    total += 8;
    for (uint64_t i = 0; i < 2 * t->layer_size[k]; ++i) total += (uint64_t)orc_ptr_bytes(t->layers[k][i]);
  }
  return total;
}

static void put_be(uint8_t* out, uint64_t v, int bytes) { /* include/utility.h:178-184 */
  for (int i = 0; i < bytes; ++i) out[i] = (uint8_t)(v >> (8 * (bytes - 1 - i)));
}

/* src/shared_tree.cpp:504-513; returns the stream length, writes if it fits */
uint64_t orc_serialize(const orc_tree* t, uint8_t* out, uint64_t cap) {
  const uint64_t total = orc_bytes(t);
  if (total > cap) return total;
  uint64_t o = 0;
  const int leaf_bytes = (t->S + 1) / 2;
  o += (uint64_t)orc_ptr_serialize(t->root, out + o);
  put_be(out + o, t->n_leaves, 8);
  o += 8;
  for (uint64_t i = 0; i < t->n_leaves; ++i, o += (uint64_t)leaf_bytes) put_be(out + o, t->leaves[i], leaf_bytes);
  for (int k = 0; k < t->n_layers; ++k) {
    put_be(out + o, t->layer_size[k], 8);
    o += 8;
    for (uint64_t i = 0; i < 2 * t->layer_size[k]; ++i) o += (uint64_t)orc_ptr_serialize(t->layers[k][i], out + o);
  }
  return o;
}

/* src/shared_tree.cpp:520-538: layers are read until the stream ends. */
orc_tree* orc_deserialize(const uint8_t* in, uint64_t len, int S) {
  orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
  t->S = S;
  uint64_t o = 0;
  const int leaf_bytes = (S + 1) / 2;
  o += (uint64_t)ptr_deserialize(in, len, &t->root);
  uint64_t n = 0;
  for (int i = 0; i < 8; ++i) n = (n << 8) | in[o + i];
  o += 8;
  t->n_leaves = n;
  t->leaves = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = 0;
    for (int b = 0; b < leaf_bytes; ++b) v = (v << 8) | in[o++];
    t->leaves[i] = v;
  }
  t->layer_size = (uint64_t*)calloc(64, sizeof(uint64_t));
  t->layers = (uint32_t**)calloc(64, sizeof(uint32_t*));
  while (o + 8 <= len && t->n_layers < 64) {
    uint64_t count = 0;
    for (int i = 0; i < 8; ++i) count = (count << 8) | in[o + i];
    o += 8;
    uint32_t* nodes = (uint32_t*)malloc((count ? count : 1) * 2 * sizeof(uint32_t));
    for (uint64_t i = 0; i < 2 * count; ++i) o += (uint64_t)ptr_deserialize(in + o, len - o, &nodes[i]);
    t->layer_size[t->n_layers] = count;
    t->layers[t->n_layers++] = nodes;
  }
  return t;
}

/* ------------------------------------------------------------------------ */
/* traversal                                                                */
/* ------------------------------------------------------------------------ */

/* src/shared_tree.cpp:252-259 (children) summed bottom-up instead of by
 * recursion; include/shared_tree.h:165 (width = children(top, root)). */
uint64_t orc_width(const orc_tree* t) {
  uint64_t* below = NULL;
  for (int k = 0; k < t->n_layers; ++k) {
    uint64_t* cnt = (uint64_t*)malloc((t->layer_size[k] ? t->layer_size[k] : 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < t->layer_size[k]; ++i) {
      uint64_t c = 0;
      for (int side = 0; side < 2; ++side) {
        const uint32_t p = t->layers[k][2 * i + side];
        if ((p & ORC_KEY31) == ORC_IDX_MASK) continue;
        c += k == 0 ? 1 : below[p & ORC_IDX_MASK];
      }
      cnt[i] = c;
    }
    free(below);
    below = cnt;
  }
  const uint64_t w = below[t->root & ORC_IDX_MASK];
  free(below);
  return w;
}

/* src/shared_tree.cpp:231-236 */
static uint64_t access_leaf(const orc_tree* t, uint32_t p) {
  uint64_t v = t->leaves[p & ORC_IDX_MASK];
  if (p & ORC_MIRROR) v = orc_mirrored(v, t->S);
  if (p & ORC_TRANSPOSE) v = orc_transposed(v);
  return v;
}

/* src/shared_tree.cpp:553-614: explicit-stack DFS.  Children continue as
 * pointer{child, top.mirror, top.transpose}; a mirrored parent visits right
 * before left; null children are skipped BEFORE composing (:606-612). */
uint64_t orc_decode(const orc_tree* t, uint64_t* out, uint64_t cap) {
  typedef struct {
    int layer; /* -1 = leaf level */
    uint32_t p;
  } frame;
  frame* stack = (frame*)malloc((size_t)(2 * t->n_layers + 4) * sizeof(frame));
  int sp = 0;
  uint64_t n = 0;
  if ((t->root & ORC_KEY31) != ORC_IDX_MASK) stack[sp++] = (frame){t->n_layers - 1, t->root};
  while (sp > 0) {
    const frame f = stack[--sp];
    if (f.layer < 0) {
      if (n < cap) out[n] = access_leaf(t, f.p);
      ++n;
      continue;
    }
    const uint32_t l = t->layers[f.layer][2 * (f.p & ORC_IDX_MASK)];
    const uint32_t r = t->layers[f.layer][2 * (f.p & ORC_IDX_MASK) + 1];
    const int m = (f.p >> 29) & 1, tr = (f.p >> 30) & 1;
    const uint32_t first = m ? r : l, second = m ? l : r;
    if ((second & ORC_KEY31) != ORC_IDX_MASK) stack[sp++] = (frame){f.layer - 1, orc_compose(second, m, tr)};
    if ((first & ORC_KEY31) != ORC_IDX_MASK) stack[sp++] = (frame){f.layer - 1, orc_compose(first, m, tr)};
  }
  free(stack);
  return n;
}

/* src/shared_tree.cpp:268-291: the left subtree of layer L holds 1<<L leaves. */
void orc_random_access(const orc_tree* t, const uint64_t* idx, uint64_t q, uint64_t* out) {
  for (uint64_t j = 0; j < q; ++j) {
    uint64_t index = idx[j];
    uint32_t cur = t->root;
    for (int layer = t->n_layers - 1; layer >= 0; --layer) {
      const uint32_t l = t->layers[layer][2 * (cur & ORC_IDX_MASK)];
      const uint32_t r = t->layers[layer][2 * (cur & ORC_IDX_MASK) + 1];
      const int m = (cur >> 29) & 1, tr = (cur >> 30) & 1;
      const uint32_t first = m ? r : l, second = m ? l : r;
      const uint64_t size = 1ull << layer;
      if (index < size) {
        cur = orc_compose(first, m, tr);
      } else {
        index -= size;
        cur = orc_compose(second, m, tr);
      }
    }
    out[j] = access_leaf(t, cur);
  }
}

/* ------------------------------------------------------------------------ */
/* synthetic genome-shaped data (this repo's workload; not in the reference) */
/* ------------------------------------------------------------------------ */

static uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* Base i of the underlying i.i.d. sequence: 2 bits of splitmix(seed, i/32). */
static int base_code(uint64_t seed, uint64_t i) {
  const uint64_t w = splitmix(seed ^ ((i >> 5) * 0xD1B54A32D192ED03ull));
  return (int)((w >> (2 * (i & 31))) & 3);
}

#define SYNTH_ALIGN (12ull * 1024ull)

/* Planted repeats: disjoint destination intervals in increasing order. Each
 * copies from the UNDERLYING sequence (so the result is a pure function of
 * the position).  Lengths 300*2^k + r (k in 0..9, r < 300*2^k) capped at
 * 200000; gaps uniform with the mean that yields the requested coverage;
 * of every 8 copies: 4 have (dst-src) a multiple of 12*1024 bases (they dedup
 * up to node layer 9), 1 is an aligned reverse-complement copy, 3 are at an
 * arbitrary distance. */
uint64_t orc_synth_repeats(uint64_t n_bases, uint64_t seed, uint32_t repeat_permille, orc_repeat* out,
                           uint64_t cap) {
  if (repeat_permille == 0 || repeat_permille >= 1000) return 0;
  uint64_t state = splitmix(seed ^ 0x5eedc0de5eedc0deull);
  uint64_t cursor = 0, count = 0;
  for (;;) {
    state = splitmix(state);
    const unsigned k = (unsigned)(state % 10);
    state = splitmix(state);
    uint64_t len = (300ull << k) + state % (300ull << k);
    if (len > 200000) len = 200000;
    /* gap uniform in [0, 2*len*(1000-p)/p): E[gap]/E[len] = (1-f)/f */
    state = splitmix(state);
    const uint64_t span = 2 * len * (1000 - repeat_permille) / repeat_permille + 1;
    const uint64_t dst = cursor + state % span;
    if (dst + len > n_bases) break;
    state = splitmix(state);
    uint64_t src = state % (n_bases - len + 1);
    state = splitmix(state);
    const unsigned cls = (unsigned)(state & 7);
    uint32_t rc = 0;
    if (cls < 4) { /* aligned forward copy */
      const uint64_t d = dst > src ? dst - src : src - dst;
      const uint64_t da = d - d % SYNTH_ALIGN;
      src = dst > src ? dst - da : (dst + da + len <= n_bases ? dst + da : dst);
    } else if (cls == 4) { /* aligned reverse-complement copy: src+len-dst = 0 mod ALIGN */
      rc = 1;
      const uint64_t want = (SYNTH_ALIGN - (src + len) % SYNTH_ALIGN + dst % SYNTH_ALIGN) % SYNTH_ALIGN;
      if (src + want + len <= n_bases) src += want;
      else if (src >= SYNTH_ALIGN - want) src -= SYNTH_ALIGN - want;
    }
    if (count < cap) out[count] = (orc_repeat){dst, src, len, rc, 0};
    ++count;
    cursor = dst + len;
  }
  return count;
}

void orc_synth_fill(char* out, uint64_t first, uint64_t count, uint64_t seed, const orc_repeat* reps,
                    uint64_t n_reps) {
  static const char acgt[4] = {'A', 'C', 'G', 'T'};
  uint64_t r = 0;
  /* first repeat whose end is beyond `first` (intervals are sorted, disjoint) */
  {
    uint64_t lo = 0, hi = n_reps;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) / 2;
      if (reps[mid].dst + reps[mid].len <= first) lo = mid + 1;
      else hi = mid;
    }
    r = lo;
  }
  for (uint64_t j = 0; j < count; ++j) {
    const uint64_t i = first + j;
    while (r < n_reps && reps[r].dst + reps[r].len <= i) ++r;
    int code;
    if (r < n_reps && i >= reps[r].dst) {
      const uint64_t off = i - reps[r].dst;
      code = reps[r].rc ? 3 - base_code(seed, reps[r].src + reps[r].len - 1 - off)
                        : base_code(seed, reps[r].src + off);
    } else {
      code = base_code(seed, i);
    }
    out[j] = acgt[code];
  }
}

/* "Real genome" variant of the synthetic workload (this repo's own definition; device twin:
 * csrc/synth.cu masked_base): runs of N and soft-masked (lower-case) stretches, pure functions of the position. */
void orc_synth_mask(char* text, uint64_t first, uint64_t count, uint64_t seed) {
  for (uint64_t j = 0; j < count; ++j) {
    const uint64_t i = first + j, b = i >> 16;
    const uint64_t h = splitmix(seed ^ 0x4e4e4e4e4e4e4e4eull ^ (b * 0x9E3779B97F4A7C15ull));
    char c = text[j];
    if (h % 100ull < 8ull) {
      const uint64_t start = (b << 16) + ((h >> 8) % 49152ull), len = 1ull + ((h >> 32) % 16384ull);
      if (i >= start && i < start + len) c = 'N';
    }
    const uint64_t h2 = splitmix(seed ^ 0x6c6f776572636173ull ^ ((i >> 12) * 0xD1B54A32D192ED03ull));
    if (h2 & 1ull) c = (char)(c | 0x20);
    text[j] = c;
  }
}
