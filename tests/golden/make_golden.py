#!/usr/bin/env python
"""Generate tests/golden/* by executing the REFERENCE's own function bodies here.

Run in the build container only (needs /root/reference; the GPU box has no such
path and never runs this):

    python tests/golden/make_golden.py

What it does: pulls individual function definitions out of the reference sources
with ``ast`` (nothing is copied into the repo; the reference packages cannot be
imported wholesale -- ``sentence_transformers`` is absent and
``easyeditor/__init__`` star-imports model zoos written for transformers 4.31),
executes them in a namespace where

  * ``sentence_transformers`` is ``oracle/st_shim.py`` (stub encoder + the oracle's
    ``util``), because the real package is not installed -- so the ARITHMETIC under
    these fixtures is the oracle's (parity unpinned, see oracle/__init__.py) while
    the GLUE is the reference's own code;
  * ``torch.tensor(...).to(0)`` / ``torch.device('cuda:..')`` are redirected to the
    CPU (no GPU in this container);
  * ``open()`` of the author's hard-coded bank paths (dataset/vllm.py:95,109) is
    redirected to the bank pickle this script has just produced with the
    reference's ``encode_ike_facts_multimodal``.

Outputs: ``evqa_bank.npz`` (bank embeddings) and ``evqa_golden.json`` (bank text
columns, queries, and the reference functions' outputs).

Level-0 call trace (``evqa_trace.json`` + ``evqa_trace.npz``): while the reference bodies run, every call they make
into ``util.*`` is recorded -- which reference function made it, the exact tensors it passed (query, bank, kwargs)
and the value it got back.  ``tests/test_gpu_parity.py::test_level0_call_trace_replay`` replays those calls
against ``devqa_b200.util`` on the GPU: the drop-in boundary exercised with the reference's own arguments, without
any reference source in the repo.  Distinct arrays are stored once (keyed by sha256).
"""
import ast
import builtins
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import semantic_search as oracle_util   # noqa: E402
from oracle import st_shim                          # noqa: E402

REF = "/root/reference/DE-VQA"
N_RECORDS = 40
N_QUERIES = 30
K_DEMOS = 16


def load_def(rel_path, name, cls=None):
    """Return the AST of function ``name`` (last definition wins, as in Python)."""
    src = open(os.path.join(REF, rel_path)).read()
    tree = ast.parse(src)
    scope = tree.body
    if cls is not None:
        scope = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][-1].body
    node = [n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name][-1]
    node.decorator_list = []
    mod = ast.Module(body=[node], type_ignores=[])
    return compile(ast.fix_missing_locations(mod), os.path.join(REF, rel_path), "exec")


class _CpuTensor(torch.Tensor):
    """``.to(<cuda thing>)`` is a no-op: the reference hard-codes ``.to(0)``."""

    def to(self, *a, **k):
        return self


class _TorchProxy(types.ModuleType):
    def __init__(self):
        super().__init__("torch")

    def __getattr__(self, item):
        return getattr(torch, item)

    @staticmethod
    def tensor(data, *a, **k):
        return torch.as_tensor(np.asarray(data) if not isinstance(data, torch.Tensor) else data.detach().clone()
                               ).as_subclass(_CpuTensor)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


class TraceUtil:
    """``util`` as the reference bodies see it: the oracle, with every call recorded."""

    def __init__(self, inner):
        self.inner = inner
        self.site = "?"
        self.calls = []
        self.arrays = {}
        self.dot_score = inner.dot_score
        self.cos_sim = inner.cos_sim

    def _key(self, t):
        import hashlib
        a = np.ascontiguousarray(torch.as_tensor(t).detach().cpu().numpy().astype(np.float32))
        k = "a" + hashlib.sha256(a.tobytes() + str(a.shape).encode()).hexdigest()[:16]
        self.arrays.setdefault(k, a)
        return k

    def normalize_embeddings(self, x):
        out = self.inner.normalize_embeddings(x)
        self.calls.append({"site": self.site, "fn": "normalize_embeddings", "input": self._key(x), "output": self._key(out)})
        return out

    def semantic_search(self, query_embeddings, corpus_embeddings, query_chunk_size=100, corpus_chunk_size=500000,
                        top_k=10, score_function=None):
        kw = {} if score_function is None else {"score_function": score_function}
        hits = self.inner.semantic_search(query_embeddings, corpus_embeddings, query_chunk_size, corpus_chunk_size,
                                          top_k, **kw)
        name = "cos_sim" if score_function is None else getattr(score_function, "__name__", "?")
        self.calls.append({"site": self.site, "fn": "semantic_search", "query": self._key(query_embeddings),
                           "corpus": self._key(corpus_embeddings), "top_k": int(top_k), "score_function": name,
                           "query_ndim": int(torch.as_tensor(query_embeddings).dim()),
                           "result": [[[h["corpus_id"], h["score"]] for h in row] for row in hits]})
        return hits


def main():
    global oracle_util
    plain_oracle = oracle_util
    tracer = oracle_util = TraceUtil(plain_oracle)          # every namespace below hands the reference THIS util
    shim = st_shim.install(oracle_util)
    tproxy = _TorchProxy()
    tmp = tempfile.mkdtemp(prefix="devqa_golden_")

    raw = json.load(open(os.path.join(REF, "data/easy-edit-mm/vqa/vqa_train.json")))[:N_RECORDS]
    evalq = json.load(open(os.path.join(REF, "data/easy-edit-mm/vqa/vqa_eval.json")))[:N_QUERIES]

    # dataset items as easyeditor/dataset/vqa.py:119-137 shapes them (text fields only)
    records = [{
        'prompt': r['src'], 'target': r['alt'], 'rephrase_prompt': r['rephrase'],
        'locality_prompt': r['loc'], 'locality_ground_truth': r['loc_ans'],
        'image_path': r['image'], 'rephrase_image_path': r['image_rephrase'],
        'locality_image_path': r['m_loc'],
    } for r in raw]

    hparams = types.SimpleNamespace(results_dir=tmp, task_name="vqa", sentence_model_name="stub/all-MiniLM-L6-v2",
                                    k=K_DEMOS, device="cpu", alg_name="IKE")
    encoder = shim.SentenceTransformer("stub")

    # ---- bank producer: easyeditor/models/ike/util.py:54-86 ------------------------------
    ns = {"tqdm": lambda it, **k: it, "os": os, "pickle": pickle, "print": lambda *a, **k: None,
          "SentenceTransformer": shim.SentenceTransformer, "Dataset": object,
          "IKEMultimodalHyperParams": object}
    exec(load_def("easyeditor/models/ike/util.py", "encode_ike_facts_multimodal"), ns)
    ns["encode_ike_facts_multimodal"](encoder, records, hparams)
    bank_path = os.path.join(tmp, "embedding", "vqa_embeddings_llava.pkl")
    bank = pickle.load(open(bank_path, "rb"))
    assert set(bank) == {"sentences", "embeddings", "images", "prompts"}
    # the older two-key layout that apply_ike_to_multimodal_model still reads (ike_main.py:188-189)
    with open(os.path.join(tmp, "embedding", "vqa_embeddings.pkl"), "wb") as f:
        pickle.dump({"sentences": bank["sentences"], "embeddings": bank["embeddings"]}, f,
                    protocol=pickle.HIGHEST_PROTOCOL)

    # ---- dataset/vllm.py: init_retrieval + finds_sim, run unmodified ----------------------
    def redirected_open(path, *a, **k):
        if isinstance(path, str) and path.endswith(("vqa_embeddings_llava.pkl", "vlkeb_embeddings_collect.pkl")):
            path = bank_path
        return builtins.open(path, *a, **k)

    ns_ds = {"torch": tproxy, "util": oracle_util, "SentenceTransformer": shim.SentenceTransformer,
             "pickle": pickle, "open": redirected_open}
    exec(load_def("dataset/vllm.py", "init_retrieval", cls="BaseVLLMEditData"), ns_ds)
    exec(load_def("dataset/vllm.py", "finds_sim", cls="BaseVLLMEditData"), ns_ds)
    self_obj = types.SimpleNamespace()
    tracer.site = "dataset/vllm.py:89 init_retrieval"
    ns_ds["init_retrieval"](self_obj, types="vqa")
    finds = []
    for r in evalq:
        for trg in (r['pred'], r['alt']):
            tracer.site = "dataset/vllm.py:65 finds_sim"
            ex, img = ns_ds["finds_sim"](self_obj, r['src'], trg)
            q = plain_oracle.normalize_embeddings(torch.tensor(encoder.encode(r['src'])).unsqueeze(0))
            hits = plain_oracle.semantic_search(q, torch.as_tensor(self_obj.stored_embeddings),
                                                score_function=plain_oracle.dot_score, top_k=5)[0]
            finds.append({"src": r['src'], "trg": trg, "example": list(ex), "image": img,
                          "hit_ids": [h["corpus_id"] for h in hits],
                          "hit_scores": [h["score"] for h in hits]})

    # ---- IKE / HICE lookups ------------------------------------------------------------
    ns_ike = {"torch": tproxy, "util": oracle_util, "SentenceTransformer": shim.SentenceTransformer,
              "pickle": pickle, "AutoModelForCausalLM": object, "AutoTokenizer": object,
              "IKEHyperParams": object, "IKEMultimodalHyperParams": object,
              "Any": object, "Dict": dict, "List": list, "Tuple": tuple}
    exec(load_def("easyeditor/models/ike/ike_main.py", "apply_ike_to_multimodal_model"), ns_ike)
    exec(load_def("easyeditor/models/ike/ike_main.py", "apply_hice_to_multimodal_model"), ns_ike)
    ns_ev = {"torch": tproxy, "util": oracle_util}
    exec(load_def("easyeditor/evaluate/multimodal_evaluate.py", "select_icl_samples"), ns_ev)

    stored = {"sentences": bank["sentences"], "embeddings": torch.tensor(bank["embeddings"])}
    lookups = []
    for r in evalq[:12]:
        req = {"prompt": r['src'], "target": r['alt']}
        tracer.site = "easyeditor/models/ike/ike_main.py:171 apply_ike_to_multimodal_model"
        ike = ns_ike["apply_ike_to_multimodal_model"](None, None, req, hparams, train_ds=records)
        tracer.site = "easyeditor/models/ike/ike_main.py:210 apply_hice_to_multimodal_model"
        hice = ns_ike["apply_hice_to_multimodal_model"](None, None, req, hparams, train_ds=records,
                                                        sentence_model=encoder, stored_data=stored)
        tracer.site = "easyeditor/evaluate/multimodal_evaluate.py:395 select_icl_samples"
        sel = ns_ev["select_icl_samples"](r['src'], r['alt'], stored, encoder, "cpu", hparams)
        hp40 = types.SimpleNamespace(**{**vars(hparams), "k": 40})   # long enough to hit the 2500-char cap
        sel40 = ns_ev["select_icl_samples"](r['src'], r['alt'], stored, encoder, "cpu", hp40)
        to_idx = {s: i for i, s in enumerate(bank["sentences"])}
        lookups.append({"prompt": r['src'], "target": r['alt'],
                        "ike_ids": [to_idx[s] for s in ike[:-1]], "ike_tail": ike[-1],
                        "hice_ids": [to_idx[s] for s in hice[:-1]], "hice_tail": hice[-1],
                        "select_ids": [to_idx[s] for s in sel],
                        "select_k40_ids": [to_idx[s] for s in sel40]})

    np.savez_compressed(os.path.join(HERE, "evqa_bank.npz"), embeddings=np.asarray(bank["embeddings"], np.float32))
    with open(os.path.join(HERE, "evqa_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "reference_functions": [
                       "easyeditor/models/ike/util.py:54 encode_ike_facts_multimodal",
                       "dataset/vllm.py:89 init_retrieval", "dataset/vllm.py:65 finds_sim",
                       "easyeditor/models/ike/ike_main.py:171 apply_ike_to_multimodal_model",
                       "easyeditor/models/ike/ike_main.py:210 apply_hice_to_multimodal_model",
                       "easyeditor/evaluate/multimodal_evaluate.py:395 select_icl_samples"],
                   "arithmetic": "oracle/semantic_search.py (sentence_transformers absent; parity unpinned)",
                   "encoder": "oracle/st_shim.py StubEncoder(dim=384, mix=0.7)",
                   "k": K_DEMOS, "records": records,
                   "bank": {"sentences": bank["sentences"], "images": bank["images"], "prompts": bank["prompts"]},
                   "finds_sim": finds, "lookups": lookups}, f, indent=1, ensure_ascii=False)
    np.savez_compressed(os.path.join(HERE, "evqa_trace.npz"), **tracer.arrays)
    with open(os.path.join(HERE, "evqa_trace.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (TraceUtil)",
                   "what": "every util.* call the reference's own function bodies made while the golden file was "
                           "generated: arguments (arrays in evqa_trace.npz, keyed by sha256) and returned value",
                   "calls": tracer.calls}, f, indent=0)
    print("trace:", len(tracer.calls), "util calls,", len(tracer.arrays), "distinct arrays")
    st_shim.uninstall()
    print("wrote", len(bank["sentences"]), "bank rows,", len(finds), "finds_sim cases,", len(lookups), "lookups")


if __name__ == "__main__":
    main()
