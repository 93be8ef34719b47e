"""Host-side mirror of ``Bm25Index`` (crates/velesdb-core/src/index/bm25.rs) over the C ABI.

The host keeps what the Rust type keeps on the host -- the tokenizer (bm25.rs:114-120), the
string -> term-id dictionary, per-document term frequencies -- and freezes them into the CSR snapshot
``veles_bm25_from_csr`` uploads.  Scoring and top-k run on the GPU (csrc/bm25.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nv


def tokenize(text: str):
    """bm25.rs:114-120: lowercase, split on non-alphanumeric chars, drop tokens of byte length <= 1."""
    out, cur = [], []
    for ch in text.lower():
        if ch.isalnum():
            cur.append(ch)
        elif cur:
            out.append("".join(cur))
            cur = []
    if cur:
        out.append("".join(cur))
    return [t for t in out if len(t.encode("utf-8")) > 1]


class Bm25Params:
    def __init__(self, k1=1.2, b=0.75):  # bm25.rs:47-58
        self.k1, self.b = k1, b


class Bm25Snapshot:
    """Owns one ``veles_bm25_t*`` built from CSR arrays."""

    def __init__(self, term_ptr, post_doc, post_tf, df, doc_len, doc_count, total_len, k1=1.2, b=0.75):
        nv.init()
        self._keep = [np.ascontiguousarray(term_ptr, np.uint64), np.ascontiguousarray(post_doc, np.uint32),
                      np.ascontiguousarray(post_tf, np.uint32), np.ascontiguousarray(df, np.uint32),
                      np.ascontiguousarray(doc_len, np.uint32)]
        tp, pd, pt, dfa, dl = self._keep
        self.n_terms = tp.size - 1
        h = C.c_void_p()
        nv.check(nv.lib().veles_bm25_from_csr(self.n_terms, nv.ptr(tp), nv.ptr(pd), nv.ptr(pt), nv.ptr(dfa), dl.size,
                                              nv.ptr(dl), int(doc_count), int(total_len), k1, b, C.byref(h)))
        self.h = h
        self._keep = None

    def __del__(self):
        if getattr(self, "h", None):
            try:
                nv.lib().veles_bm25_free(self.h)
            except Exception:
                pass
            self.h = None

    def search_batch(self, q_ptr, q_terms, k, stream=None):
        q_ptr = np.ascontiguousarray(q_ptr, np.uint32)
        q_terms = np.ascontiguousarray(q_terms, np.uint32)
        nq = q_ptr.size - 1
        docs = np.empty((nq, k), np.uint32)
        sc = np.empty((nq, k), np.float32)
        cnt = np.zeros(nq, np.uint32)
        qt = q_terms if q_terms.size else np.zeros(1, np.uint32)
        nv.check(nv.lib().veles_bm25_search_batch(self.h, nv.ptr(q_ptr), nv.ptr(qt), nq, k, nv.ptr(docs), nv.ptr(sc),
                                                  nv.ptr(cnt), stream))
        return docs, sc, cnt


class Bm25Index:
    """``Bm25Index`` (bm25.rs:78-90).  Mutations are staged on the host; the device snapshot is rebuilt
    on the next search."""

    def __init__(self, params=None):
        self.params = params or Bm25Params()
        self.vocab = {}
        self._docs = {}       # id -> {term_id: tf}
        self._doc_len = {}    # id -> token count
        self._df = {}         # term_id -> PostingList::len (stale postings of replaced docs stay, bm25.rs:188-196)
        self._posting_sets = {}
        self._total = 0
        self._snap = None

    @classmethod
    def new(cls):
        return cls()

    @classmethod
    def with_params(cls, params):
        return cls(params)

    def _term(self, tok, grow):
        t = self.vocab.get(tok)
        if t is None and grow:
            t = self.vocab[tok] = len(self.vocab)
        return t

    def add_document(self, id, text):
        if not (0 <= id <= 0xFFFFFFFF):
            raise ValueError(f"BM25 document ID {id} exceeds u32::MAX ({0xFFFFFFFF}). "
                             "The BM25 index uses RoaringBitmap which only supports 32-bit IDs.")
        toks = tokenize(text)
        if not toks:
            return
        tf = {}
        for t in toks:
            ti = self._term(t, True)
            tf[ti] = tf.get(ti, 0) + 1
        for ti in tf:
            self._posting_sets.setdefault(ti, set()).add(id)
        if id in self._docs:
            self._total = max(0, self._total - self._doc_len[id])
        self._docs[id] = tf
        self._doc_len[id] = len(toks)
        self._total += len(toks)
        self._snap = None

    def remove_document(self, id) -> bool:
        if not (0 <= id <= 0xFFFFFFFF):
            raise ValueError(f"BM25 document ID {id} exceeds u32::MAX")
        tf = self._docs.pop(id, None)
        if tf is None:
            return False
        for ti in tf:
            s = self._posting_sets.get(ti)
            if s is not None:
                s.discard(id)
                if not s:
                    del self._posting_sets[ti]
        self._total = max(0, self._total - self._doc_len.pop(id))
        self._snap = None
        return True

    def len(self):
        return len(self._docs)

    __len__ = len

    def is_empty(self):
        return not self._docs

    def term_count(self):
        return len(self._posting_sets)

    def _snapshot(self):
        if self._snap is None:
            # External document ids (any u32; the reference keys hash maps by them, bm25.rs:62-90) are remapped to dense
            # slots in ascending-id order, so one add_document(4_000_000_000, ..) costs one slot, not 16 GB of
            # doc_len, and "ties by ascending id" is still "ties by ascending slot".  Results are mapped back.
            ids = sorted(self._docs)
            self._slot_ids = np.array(ids or [0], np.uint32)
            slot_of = {d: i for i, d in enumerate(ids)}
            n_terms = len(self.vocab)
            lists = [[] for _ in range(n_terms)]
            for d in ids:
                for ti, f in self._docs[d].items():
                    lists[ti].append((slot_of[d], f))
            term_ptr = np.zeros(n_terms + 1, np.uint64)
            for t in range(n_terms):
                term_ptr[t + 1] = term_ptr[t] + len(lists[t])
            flat = [p for l in lists for p in l]
            post_doc = np.array([p[0] for p in flat] or [0], np.uint32)
            post_tf = np.array([p[1] for p in flat] or [0], np.uint32)
            df = np.array([len(self._posting_sets.get(t, ())) for t in range(n_terms)] or [0], np.uint32)
            slots = len(ids)
            doc_len = np.zeros(max(slots, 1), np.uint32)
            for d, l in self._doc_len.items():
                if d in slot_of:
                    doc_len[slot_of[d]] = l
            self._snap = Bm25Snapshot(term_ptr, post_doc[:len(flat)] if flat else post_doc, post_tf, df[:max(n_terms, 1)],
                                      doc_len[:max(slots, 1)], len(self._docs), self._total, self.params.k1, self.params.b)
        return self._snap

    def search(self, query, k):
        return self.search_batch([query], k)[0]

    def search_batch(self, queries, k):
        toks = [tokenize(q) for q in queries]
        if not self._docs:
            return [[] for _ in queries]
        q_ptr = np.zeros(len(queries) + 1, np.uint32)
        terms = []
        for i, ts in enumerate(toks):
            for t in ts:
                ti = self._term(t, False)
                terms.append(nv.INVALID_ID if ti is None else ti)
            q_ptr[i + 1] = len(terms)
        docs, sc, cnt = self._snapshot().search_batch(q_ptr, np.array(terms or [0], np.uint32)[:len(terms)], k)
        ext = self._slot_ids
        return [[(int(ext[docs[i, j]]), float(sc[i, j])) for j in range(int(cnt[i]))] for i in range(len(queries))]
