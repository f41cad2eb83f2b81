"""Host-side mirror of the reference glue (de-vqa_b200/retrieval.py), the util shim's argument
handling and the sharding arithmetic, checked on the CPU.  The search itself is GPU-only in the
product; here a stand-in bank backed by the ORACLE plays the device so that the host logic can be
compared with the golden outputs of the reference's own functions."""
import os
import pickle
import types

import numpy as np
import pytest
import torch

import devqa_b200
from devqa_b200 import retrieval, dist as ddist
from devqa_b200.bank import hits_from_arrays, norm_flags
from oracle import semantic_search as ss
from oracle.st_shim import StubEncoder


class OracleBank:
    """FactBank look-alike whose search is the CPU oracle (tests only)."""

    def __init__(self, emb):
        self.emb = torch.as_tensor(np.asarray(emb)).float()
        self.sentences, self.images, self.prompts = [], [], []

    @classmethod
    def from_stored_data(cls, stored, device=0, dtype="fp32"):
        b = cls(stored["embeddings"])
        b.sentences = list(stored.get("sentences", []))
        b.images = list(stored.get("images", []))
        b.prompts = [list(p) for p in stored.get("prompts", [])]
        return b

    @classmethod
    def from_embeddings(cls, emb, device=0, dtype="fp32", capacity=0):
        return cls(emb)

    @classmethod
    def from_pickle(cls, path, device=0, dtype="fp32"):
        with open(path, "rb") as f:
            return cls.from_stored_data(pickle.load(f))

    def auto_shadow(self):
        return False

    def __len__(self):
        return len(self.emb)

    def close(self):
        pass

    def search_host(self, q, k, normalize_bank="eps", normalize_query="eps", flags=0):
        m = lambda v: "none" if v in (None, False) else ("eps" if v in (True, "eps") else "noeps")
        s, i = ss.search_arrays(torch.as_tensor(np.asarray(q)).float(), self.emb, k, m(normalize_bank), m(normalize_query))
        pad = k - s.shape[1]
        if pad > 0:
            s = np.pad(s, ((0, 0), (0, pad)), constant_values=-np.inf)
            i = np.pad(i, ((0, 0), (0, pad)), constant_values=-1)
        return s, i

    def hits(self, q, k, normalize_bank="eps", normalize_query="eps"):
        return hits_from_arrays(*self.search_host(q, k, normalize_bank, normalize_query))


@pytest.fixture()
def oracle_backed(monkeypatch):
    retrieval.clear_bank_cache()
    monkeypatch.setattr(retrieval, "FactBank", OracleBank)
    yield
    retrieval._file_banks.clear()
    retrieval._object_banks.clear()


def _write_bank(tmp_path, golden, two_key=False):
    stored = {"sentences": golden["bank"]["sentences"], "embeddings": golden["embeddings"]}
    if not two_key:
        stored.update(images=golden["bank"]["images"], prompts=golden["bank"]["prompts"])
    os.makedirs(tmp_path / "embedding", exist_ok=True)
    path = tmp_path / "embedding" / ("vqa_embeddings.pkl" if two_key else "vqa_embeddings_llava.pkl")
    with open(path, "wb") as f:
        pickle.dump(stored, f, protocol=pickle.HIGHEST_PROTOCOL)
    return str(path)


def test_encode_ike_facts_multimodal_layout(tmp_path, golden):
    hp = types.SimpleNamespace(results_dir=str(tmp_path), task_name="vqa", sentence_model_name="x/stub")
    stored = retrieval.encode_ike_facts_multimodal(StubEncoder(), golden["records"], hp)
    on_disk = pickle.load(open(tmp_path / "embedding" / "vqa_embeddings_llava.pkl", "rb"))
    assert set(on_disk) == {"sentences", "embeddings", "images", "prompts"}
    assert on_disk["sentences"] == golden["bank"]["sentences"] == stored["sentences"]
    assert on_disk["images"] == golden["bank"]["images"] and on_disk["prompts"] == golden["bank"]["prompts"]
    assert on_disk["embeddings"].dtype == np.float32
    np.testing.assert_array_equal(on_disk["embeddings"], golden["embeddings"])


def test_encode_ike_facts_text_layout(tmp_path, golden):
    class DS(list):
        pass
    ds = DS(dict(r, target_new=r["target"]) for r in golden["records"][:5])
    hp = types.SimpleNamespace(results_dir=str(tmp_path), alg_name="IKE", sentence_model_name="a/b/minilm")
    retrieval.encode_ike_facts(StubEncoder(), ds, hp)
    got = pickle.load(open(tmp_path / "IKE" / "embedding" / "minilm_DS_5.pkl", "rb"))
    assert set(got) == {"sentences", "embeddings"} and got["sentences"] == golden["bank"]["sentences"][:15]


def test_finds_sim_matches_reference(tmp_path, golden, oracle_backed):
    rd = retrieval.RetrievalData()
    rd.init_retrieval(types="vqa", bank_path=_write_bank(tmp_path, golden), sentence_model=StubEncoder())
    assert rd.stored_sentences == golden["bank"]["sentences"]
    for case in golden["finds_sim"]:
        ex, img = rd.finds_sim(case["src"], case["trg"])
        assert list(ex) == case["example"] and img == case["image"]
        ex2, img2 = rd.finds_sim_easyedit(case["src"], case["trg"])
        assert ex2 == [case["example"]] and img2 == case["image"]


def test_finds_sim_falls_back_to_last_hit(tmp_path, golden, oracle_backed):
    """dataset/vllm.py:79-81: when every hit carries the excluded target, the LAST hit is used."""
    rd = retrieval.RetrievalData()
    rd.init_retrieval(types="VLKEB", bank_path=_write_bank(tmp_path, golden), sentence_model=StubEncoder())
    rd.prompts = [[p[0], "same"] for p in rd.prompts]
    case = golden["finds_sim"][0]
    ex, img = rd.finds_sim(case["src"], "same")
    assert img == golden["bank"]["images"][case["hit_ids"][-1]]


def test_ike_and_hice_lookups_match_reference(tmp_path, golden, oracle_backed):
    _write_bank(tmp_path, golden, two_key=True)
    hp = types.SimpleNamespace(results_dir=str(tmp_path), task_name="vqa", sentence_model_name="x/stub",
                               k=golden["k"], device=0, alg_name="IKE")
    enc = StubEncoder()
    sents = golden["bank"]["sentences"]
    stored = {"sentences": sents, "embeddings": torch.tensor(golden["embeddings"])}
    for case in golden["lookups"]:
        req = {"prompt": case["prompt"], "target": case["target"]}
        ike = retrieval.apply_ike_to_multimodal_model(None, None, req, hp, train_ds=[0], sentence_model=enc)
        assert ike[:-1] == [sents[i] for i in case["ike_ids"]] and ike[-1] == case["ike_tail"]
        hice = retrieval.apply_hice_to_multimodal_model(None, None, req, hp, train_ds=[0], sentence_model=enc,
                                                        stored_data=stored)
        assert hice[:-1] == [sents[i] for i in case["hice_ids"]] and hice[-1] == case["hice_tail"]
        sel = retrieval.select_icl_samples(case["prompt"], case["target"], stored, enc, "cuda:0", hp)
        assert sel == [sents[i] for i in case["select_ids"]]
        hp40 = types.SimpleNamespace(**{**vars(hp), "k": 40})
        sel40 = retrieval.select_icl_samples(case["prompt"], case["target"], stored, enc, "cuda:0", hp40)
        assert sel40 == [sents[i] for i in case["select_k40_ids"]]


def test_select_icl_samples_with_given_query_embedding(golden, oracle_backed):
    enc = StubEncoder()
    sents = golden["bank"]["sentences"]
    stored = {"sentences": sents, "embeddings": torch.tensor(golden["embeddings"])}
    hp = types.SimpleNamespace(k=4, device=0)
    case = golden["lookups"][0]
    q = torch.tensor(enc.encode(f"New Fact: {case['prompt']} {case['target']}\nPrompt: {case['prompt']} {case['target']}\n\n")).unsqueeze(0)
    got = retrieval.select_icl_samples(case["prompt"], case["target"], stored, enc, "cuda:0", hp, query_embedding=q * 3.0)
    assert got == [sents[i] for i in case["select_ids"][:4]]      # scaling the query does not change the ranking


def test_bank_cache_reuses_and_invalidates(golden, oracle_backed):
    emb = torch.tensor(golden["embeddings"])
    stored = {"sentences": golden["bank"]["sentences"], "embeddings": emb}
    b1 = retrieval.bank_from_stored_data(stored)
    assert retrieval.bank_from_stored_data(stored) is b1            # resident: no re-upload per request
    emb.mul_(1.0)                                                   # in-place edit bumps the tensor version
    assert retrieval.bank_from_stored_data(stored) is not b1


def test_bank_cache_never_serves_a_stale_ndarray(golden, oracle_backed):
    """ndarrays have no version counter, so an equal-length in-place edit must not be served from a stale
    resident bank; pin_stored_data opts in to reuse."""
    emb = np.array(golden["embeddings"], copy=True)
    stored = {"sentences": golden["bank"]["sentences"], "embeddings": emb}
    b1 = retrieval.bank_from_stored_data(stored)
    emb[0] *= -1.0
    b2 = retrieval.bank_from_stored_data(stored)
    assert b2 is not b1 and float(b2.emb[0, 0]) == float(emb[0, 0])
    p1 = retrieval.pin_stored_data(stored)
    assert retrieval.bank_from_stored_data(stored) is p1
    assert retrieval.pin_stored_data(stored) is not p1


def test_device_spellings_agree():
    """'cuda:1'.index is a bound method (truthy); every cache key goes through _device_index."""
    import torch as _t
    assert retrieval._device_index("cuda:1") == 1 and retrieval._device_index(1) == 1
    assert retrieval._device_index(_t.device("cuda", 1)) == 1 and retrieval._device_index("cuda") == 0


def test_dead_owners_are_evicted(golden, oracle_backed):
    emb = torch.tensor(golden["embeddings"])
    retrieval.bank_from_stored_data({"sentences": golden["bank"]["sentences"], "embeddings": emb})
    assert len(retrieval._object_banks) == 1
    del emb
    import gc; gc.collect()
    retrieval.bank_from_stored_data({"sentences": golden["bank"]["sentences"], "embeddings": torch.tensor(golden["embeddings"])})
    assert len(retrieval._object_banks) == 1


def test_apply_ike_requires_train_ds(golden):
    with pytest.raises(AssertionError):
        retrieval.apply_ike_to_multimodal_model(None, None, {"prompt": "a", "target": "b"}, types.SimpleNamespace())


def test_hparams_from_yaml(tmp_path):
    p = tmp_path / "ike.yaml"
    p.write_text("alg_name: IKE\nk: 16\nresults_dir: ./results\ndevice: 0\nmodel_name: llava\n"
                 "sentence_model_name: all-MiniLM-L6-v2\ntask_name: vqa\nname: x\ncoco_image: /data\n")
    hp = retrieval.IKEMultimodalHyperParams.from_hparams(str(p).replace(".yaml", ""))
    assert hp.k == 16 and hp.task_name == "vqa" and hp.extra == {"coco_image": "/data"}
    p.write_text("alg_name: MEND\nk: 1\n")
    with pytest.raises(ValueError):
        retrieval.IKEHyperParams.from_hparams(str(p))


def test_norm_flags_and_hit_records():
    L = devqa_b200._lib
    assert norm_flags("eps", "eps") == L.NORMALIZE_BANK | L.NORMALIZE_QUERY
    assert norm_flags("noeps", None) == L.NORMALIZE_BANK | L.NOEPS
    assert norm_flags(False, False) == 0 and norm_flags(True, None) == L.NORMALIZE_BANK
    with pytest.raises(ValueError):
        norm_flags("eps", "noeps")
    with pytest.raises(ValueError):
        norm_flags("l1", None)
    hits = hits_from_arrays(np.array([[0.9, 0.5, -np.inf]], np.float32), np.array([[7, 3, -1]]))
    assert hits == [[{"corpus_id": 7, "score": pytest.approx(0.9)}, {"corpus_id": 3, "score": 0.5}]]
    assert isinstance(hits[0][0]["corpus_id"], int) and isinstance(hits[0][0]["score"], float)


def test_util_shim_rejects_unknown_score_function_and_cpu_only_box():
    u = devqa_b200.util
    with pytest.raises(ValueError):
        u.semantic_search_arrays(torch.zeros(1, 4), torch.zeros(3, 4), 2, score_function=lambda a, b: a @ b.T)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            u.semantic_search(torch.zeros(1, 4), torch.zeros(3, 4), top_k=2, score_function=u.dot_score)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            u.normalize_embeddings(torch.ones(2, 4))
    # the dense helpers keep upstream semantics
    a, b = torch.randn(2, 8), torch.randn(5, 8)
    np.testing.assert_allclose(u.dot_score(a, b).numpy(), ss.dot_score(a, b).numpy(), rtol=1e-6)
    np.testing.assert_allclose(u.cos_sim(a, b).numpy(), ss.cos_sim(a, b).numpy(), rtol=1e-6)


def test_factbank_needs_cuda():
    with pytest.raises(ValueError):
        devqa_b200.FactBank(8, device="cpu")


# ---- sharding arithmetic ------------------------------------------------------------------------------
@pytest.mark.parametrize("n,world", [(0, 2), (1, 8), (10, 3), (1_000_000, 8), (10_000_019, 8)])
def test_shard_bounds_cover_everything(n, world):
    b = ddist.shard_bounds(n, world)
    assert b[0] == 0 and b[-1] == n and len(b) == world + 1
    sizes = np.diff(b)
    assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


@pytest.mark.parametrize("chunk,world", [(1, 2), (4, 3), (65536, 8)])
def test_chunked_id_map_is_a_bijection(chunk, world):
    n = chunk * world * 3 + chunk // 2 + 1
    seen = {}
    for g in range(0, n, max(1, n // 5000)):
        r, l = ddist.global_to_local(g, chunk, world)
        assert ddist.chunk_owner(g, chunk, world) == r
        assert ddist.local_to_global(l, chunk, r, world) == g
        assert (r, l) not in seen
        seen[(r, l)] = g


def test_split_append_partitions_batches():
    chunk, world = 8, 3
    total = 0
    local_sizes = [0] * world
    for batch in [1, 1, 5, 8, 30, 2, 17, 64]:
        covered = []
        for r in range(world):
            for off, cnt in ddist.split_append(total, batch, chunk, r, world):
                covered += list(range(off, off + cnt))
                for j in range(cnt):     # rows land at consecutive local slots, and map back to their global id
                    assert ddist.local_to_global(local_sizes[r] + j, chunk, r, world) == total + off + j
                local_sizes[r] += cnt
        assert sorted(covered) == list(range(batch))
        total += batch


def test_level0_shadow_policy(monkeypatch):
    """util._shadow_wanted: a caller-owned corpus gets a 2-byte shadow only if the scans that would use it can occur
    (>= 148 tiles of 128 rows), it fits in a quarter of the free device memory, and DEVQA_LEVEL0_SHADOW is not 0.
    (Pure host logic: no GPU needed.)"""
    from types import SimpleNamespace
    from devqa_b200 import util as dutil
    free = [64 << 30]
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda dev=None: (free[0], 180 << 30))
    big = SimpleNamespace(shape=(10_000_000, 768), device="cuda:0")
    small = SimpleNamespace(shape=(148 * 128 - 1, 768), device="cuda:0")
    edge = SimpleNamespace(shape=(148 * 128, 100), device="cuda:0")
    monkeypatch.delenv("DEVQA_LEVEL0_SHADOW", raising=False)
    assert dutil._shadow_wanted(big) and dutil._shadow_wanted(edge) and not dutil._shadow_wanted(small)
    free[0] = 40 << 30                                        # 15.4 GB of shadow > a quarter of 40 GB
    assert not dutil._shadow_wanted(big) and dutil._shadow_wanted(edge)
    free[0] = 64 << 30
    monkeypatch.setenv("DEVQA_LEVEL0_SHADOW", "0")
    assert not dutil._shadow_wanted(big)
