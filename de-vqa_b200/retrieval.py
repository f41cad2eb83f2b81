"""Host-side mirror of the reference's retrieval glue, over the resident FactBank.

Same names, argument meaning and return values as the reference functions, so a DE-VQA user
can switch call sites one-for-one:

  reference                                                    here
  -----------------------------------------------------------  ---------------------------
  encode_ike_facts / encode_ike_facts_multimodal               same names
      easyeditor/models/ike/util.py:9,54
  BaseVLLMEditData.init_retrieval / finds_sim                  RetrievalData.init_retrieval /
      dataset/vllm.py:89,65                                        .finds_sim
  finds_sim of easyeditor/dataset/vqa.py:173 (list not unwrapped)  .finds_sim_easyedit
  apply_ike_to_model / apply_ike_to_multimodal_model           same names
      easyeditor/models/ike/ike_main.py:14,171
  apply_hice_to_multimodal_model   ike_main.py:210             same name
  select_icl_samples   easyeditor/evaluate/multimodal_evaluate.py:395   same name
  IKEHyperParams / IKEMultimodalHyperParams   ike_hparams.py:8,36       same names (retrieval knobs)

What changes underneath: the bank is uploaded ONCE into a FactBank and kept resident (the
reference re-reads the pickle, re-uploads and re-normalises the whole bank on every request,
ike_main.py:185-194, and on every probe, multimodal_evaluate.py:408-410); normalisation is
fused into the search kernel; the query goes to the device inside one C call.
The sentence encoder stays the caller's (``SentenceTransformer``-like object with
``.encode``): it is outside the hot path.
"""
from __future__ import annotations

import os
import pickle
import weakref
from dataclasses import dataclass, field, fields
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .bank import FactBank

DEMO_CHAR_CAP = 2500      # multimodal_evaluate.py:418


# ---- hparams: the retrieval knobs of ike_hparams.py -------------------------------
@dataclass
class IKEHyperParams:
    k: int
    results_dir: str
    device: int
    alg_name: str
    model_name: str
    sentence_model_name: str
    model_parallel: bool = False
    extra: Dict[str, Any] = field(default_factory=dict)

    @classmethod
    def from_hparams(cls, hparams_name_or_path: str):
        import yaml
        if '.yaml' not in hparams_name_or_path:
            hparams_name_or_path = hparams_name_or_path + '.yaml'
        with open(hparams_name_or_path, "r") as stream:
            config = yaml.safe_load(stream)
        if not config or config.get('alg_name') != 'IKE':
            raise ValueError(f"{cls.__name__} can not load from {hparams_name_or_path}, "
                             f"alg_name is {config.get('alg_name') if config else None}")
        known = {f.name for f in fields(cls)} - {"extra"}
        return cls(**{k: v for k, v in config.items() if k in known},
                   extra={k: v for k, v in config.items() if k not in known})


@dataclass
class IKEMultimodalHyperParams(IKEHyperParams):
    task_name: str = ""
    name: str = ""


# ---- bank producer (util.py:9-28, 54-86) -----------------------------------------------------------------------
def _fact_sentences(rec, target_key: str) -> List[str]:
    new_fact = rec['prompt'] + ' ' + rec[target_key]
    target_new = rec[target_key]
    return [f"New Fact: {new_fact}\nPrompt: {new_fact}\n\n",
            f"New Fact: {new_fact}\nPrompt: {rec['rephrase_prompt']} {target_new}\n\n",
            f"New Fact: {new_fact}\nPrompt: {rec['locality_prompt']} {rec['locality_ground_truth']}\n\n"]


def encode_ike_facts(sentence_model, ds, hparams) -> Dict:
    """Text-LLM bank (util.py:9-28): ``{sentences, embeddings}`` pickled to
    ``{results_dir}/{alg_name}/embedding/{model}_{DatasetClass}_{len}.pkl``."""
    sentences: List[str] = []
    for rec in ds:
        sentences += _fact_sentences(rec, 'target_new')
    stored = {'sentences': sentences, 'embeddings': np.asarray(sentence_model.encode(sentences), dtype=np.float32)}
    base_path = f'{hparams.results_dir}/{hparams.alg_name}/embedding'
    os.makedirs(base_path, exist_ok=True)
    safe_model_name = hparams.sentence_model_name.rsplit('/', 1)[-1]
    with open(f'{base_path}/{safe_model_name}_{type(ds).__name__}_{len(ds)}.pkl', "wb") as f:
        pickle.dump(stored, f, protocol=pickle.HIGHEST_PROTOCOL)
    return stored


def encode_ike_facts_multimodal(sentence_model, ds, hparams, device=None) -> Dict:
    """Multimodal bank in the frozen four-column layout (util.py:54-86): three rows per record
    (edit, rephrase, locality), pickled to ``{results_dir}/embedding/{task_name}_embeddings_llava.pkl``.

    ``device`` (extension): also leave the bank RESIDENT on that GPU and register it under the file
    just written, so that the first ``apply_ike_to_multimodal_model`` / ``bank_from_file`` after the build does
    not go pickle -> numpy -> torch.tensor -> .to(device) again (ike_main.py:187-194 does that on every request)."""
    sentences, images, prompts = [], [], []
    for rec in ds:
        images += [rec['image_path'], rec['rephrase_image_path'], rec['locality_image_path']]
        prompts += [[rec['prompt'], rec['target']], [rec['rephrase_prompt'], rec['target']],
                    [rec['locality_prompt'], rec['locality_ground_truth']]]
        sentences += _fact_sentences(rec, 'target')
    stored = {'sentences': sentences, 'embeddings': np.asarray(sentence_model.encode(sentences), dtype=np.float32),
              'images': images, 'prompts': prompts}
    base_path = f'{hparams.results_dir}/embedding'
    os.makedirs(base_path, exist_ok=True)
    path = f'{base_path}/{hparams.task_name}_embeddings_llava.pkl'
    with open(path, "wb") as f:
        pickle.dump(stored, f, protocol=pickle.HIGHEST_PROTOCOL)
    if device is not None:
        register_bank_for_file(path, FactBank.from_stored_data(stored, device=_device_index(device)), device)
    return stored


# ---- resident-bank caches ------------------------------------------------------------------------------------
_file_banks: Dict[Tuple[str, int, float, int], FactBank] = {}
_object_banks: Dict[int, Tuple[Any, Any, FactBank]] = {}


def clear_bank_cache() -> None:
    for b in list(_file_banks.values()) + [v[2] for v in _object_banks.values()]:
        b.close()
    _file_banks.clear()
    _object_banks.clear()


def register_bank_for_file(path: str, bank: FactBank, device=0) -> None:
    """Make ``bank`` the resident copy of the bank file at ``path`` (as it is on disk right now)."""
    st = os.stat(path)
    _file_banks[(os.path.abspath(path), st.st_size, st.st_mtime, _device_index(device))] = bank


def bank_from_file(path: str, device=0) -> FactBank:
    """Pickle/.pth -> resident FactBank, uploaded once per (file version, device)."""
    st = os.stat(path)
    dev = _device_index(device)
    key = (os.path.abspath(path), st.st_size, st.st_mtime, dev)
    bank = _file_banks.get(key)
    if bank is None:
        bank = _file_banks[key] = FactBank.from_pickle(path, device=dev)
    return bank


def bank_from_stored_data(stored_data: Dict, device=0) -> FactBank:
    """``stored_data = {'sentences', 'embeddings'}`` (multimodal_editor.py:1129-1130) -> resident
    FactBank.  torch tensors carry a version counter that every in-place op bumps, so their resident copy is
    re-uploaded exactly when the object is replaced or modified.  ndarrays / lists have no such counter: an in-place
    edit of equal length would be invisible, so they are uploaded on every call (one H2D copy -- still without the
    reference's per-call renormalisation) unless the caller registered the bank with ``pin_stored_data``."""
    emb = stored_data['embeddings']
    version = getattr(emb, "_version", None)
    dev = _device_index(device)
    for key in [k for k, v in _object_banks.items() if v[0]() is None]:       # owners that died: free their banks
        _object_banks.pop(key)[2].close()
    hit = _object_banks.get((id(emb), dev))
    pinned = hit is not None and hit[1] == "pinned"
    if hit is not None and hit[0]() is emb and (pinned or (version is not None and hit[1] == version)) \
            and len(hit[2]) == len(emb):
        bank = hit[2]
    else:
        bank = FactBank.from_embeddings(emb, device=dev)
        if version is not None:
            try:
                _object_banks[(id(emb), dev)] = (weakref.ref(emb), version, bank)
                bank.auto_shadow()                                             # (a bank that lives on: worth its shadow)
            except TypeError:
                pass
    bank.sentences = stored_data['sentences']
    return bank


def pin_stored_data(stored_data: Dict, device=0) -> FactBank:
    """Declare that ``stored_data['embeddings']`` (an ndarray: no version counter) will not be modified in place,
    so that its resident copy may be reused across calls.  Call again after editing it."""
    emb = stored_data['embeddings']
    dev = _device_index(device)
    old = _object_banks.pop((id(emb), dev), None)
    if old is not None:
        old[2].close()
    bank = FactBank.from_embeddings(emb, device=dev)
    bank.auto_shadow()
    _object_banks[(id(emb), dev)] = (weakref.ref(emb), "pinned", bank)
    bank.sentences = stored_data['sentences']
    return bank


def _device_index(device) -> int:
    if isinstance(device, int):
        return device
    if isinstance(device, str):
        return int(device.split(':')[1]) if ':' in device else 0
    return getattr(device, "index", None) or 0


# ---- dataset/vllm.py: init_retrieval + finds_sim ---------------------------------------------------------------
class RetrievalData:
    """The retrieval state ``BaseVLLMEditData`` keeps (dataset/vllm.py:89-117): encoder, text
    columns, and the bank -- here a resident FactBank instead of a normalised torch tensor."""

    def init_retrieval(self, types: str = 'VLKEB', bank_path: Optional[str] = None, sentence_model=None,
                       device=0) -> None:
        """``types`` keeps the reference's meaning ('VLKEB' in types -> the VLKEB bank, else the
        EVQA one); the two hard-coded author paths (dataset/vllm.py:92,95,107,109) become arguments."""
        if sentence_model is None:
            raise ValueError("pass sentence_model= (a SentenceTransformer-like encoder)")
        if bank_path is None:
            raise ValueError("pass bank_path= (vlkeb_embeddings_collect.pkl / vqa_embeddings_llava.pkl)")
        self.retrieval_types = types
        self.sentence_model = sentence_model
        self.bank = bank_from_file(bank_path, device=device)
        self.stored_sentences = self.bank.sentences
        self.save_image_path = self.bank.images
        self.prompts = self.bank.prompts

    def _top_hits(self, src: str, tops: int) -> List[dict]:
        q = np.asarray(self.sentence_model.encode(src, show_progress_bar=False), dtype=np.float32)
        return self.bank.hits(q[None, :], tops, normalize_bank="eps", normalize_query="eps")[0]

    def finds_sim(self, src: str, trg: str, tops: int = 5):
        """dataset/vllm.py:65-87 -> ``([prompt, target], image_path)``: the nearest of the top
        ``tops`` facts whose stored target differs from ``trg``, else the last hit."""
        hit = self._top_hits(src, tops)
        for h in hit:
            if self.prompts[h["corpus_id"]][1] != trg:
                return self.prompts[h["corpus_id"]], self.save_image_path[h["corpus_id"]]
        return self.prompts[hit[-1]["corpus_id"]], self.save_image_path[hit[-1]["corpus_id"]]

    def finds_sim_easyedit(self, src: str, trg: str, tops: int = 5):
        """easyeditor/dataset/vqa.py:173-202 variant -> ``([[prompt, target]], image_path)``."""
        ex, img = self.finds_sim(src, trg, tops)
        return [ex], img


# ---- IKE / HICE demonstration lookups ----------------------------------------------------------------------------
def _query_sentence(new_fact: str, prompt_part: str) -> str:
    return f"New Fact: {new_fact}\nPrompt: {prompt_part}\n\n"


def _encode_query(sentence_model, sentence: str, device=None) -> np.ndarray:
    try:
        emb = sentence_model.encode(sentence, show_progress_bar=False, device=device)
    except TypeError:
        emb = sentence_model.encode(sentence, show_progress_bar=False)
    if hasattr(emb, "detach"):
        emb = emb.detach().cpu().numpy()
    return np.asarray(emb, dtype=np.float32).reshape(1, -1)


def _make_encoder(hparams, kwargs):
    enc = kwargs.get('sentence_model')
    if enc is not None:
        return enc
    try:
        from sentence_transformers import SentenceTransformer
    except ImportError as e:
        raise RuntimeError("sentence_transformers is not installed: pass sentence_model=<encoder>") from e
    return SentenceTransformer(hparams.sentence_model_name, device=f'cuda:{hparams.device}')


def apply_ike_to_model(model, tok, request: Dict, hparams, copy=False, return_orig_weights=False,
                       keep_original_weight=False, train_ds=None, **kwargs: Any) -> List[str]:
    """ike_main.py:14-50.  Bank file ``{results_dir}/{alg_name}/embedding/{model}_{DS}_{len}.pkl``;
    query template ``Prompt: {prompt}``; returns the k demos + the new fact."""
    assert train_ds is not None
    sentence_model = _make_encoder(hparams, kwargs)
    safe_model_name = hparams.sentence_model_name.rsplit('/', 1)[-1]
    bank = bank_from_file(f'{hparams.results_dir}/{hparams.alg_name}/embedding/'
                          f'{safe_model_name}_{type(train_ds).__name__}_{len(train_ds)}.pkl', device=hparams.device)
    new_fact = request['prompt'] + ' ' + request['target_new']
    q = _encode_query(sentence_model, _query_sentence(new_fact, request['prompt']))
    hit = bank.hits(q, hparams.k, normalize_bank="eps", normalize_query="eps")[0]
    icl_examples = [bank.sentences[h["corpus_id"]] for h in hit]
    icl_examples.append(_query_sentence(new_fact, new_fact))
    return icl_examples


def apply_ike_to_multimodal_model(model, tok, request: Dict, hparams, copy=False, return_orig_weights=False,
                                  keep_original_weight=False, train_ds=None, **kwargs: Any) -> List[str]:
    """ike_main.py:171-208.  Bank file ``{results_dir}/embedding/{task_name}_embeddings.pkl``;
    eps-normalised bank and query; returns the k demos + the new fact."""
    assert train_ds is not None
    sentence_model = _make_encoder(hparams, kwargs)
    bank = bank_from_file(f'{hparams.results_dir}/embedding/{hparams.task_name}_embeddings.pkl',
                          device=hparams.device)
    new_fact = request['prompt'] + ' ' + request['target']
    query_sentence = _query_sentence(new_fact, new_fact)
    q = _encode_query(sentence_model, query_sentence)
    hit = bank.hits(q, hparams.k, normalize_bank="eps", normalize_query="eps")[0]
    icl_examples = [bank.sentences[h["corpus_id"]] for h in hit]
    icl_examples.append(query_sentence)
    return icl_examples


def apply_hice_to_multimodal_model(model, tok, request: Dict, hparams, copy=False, return_orig_weights=False,
                                   keep_original_weight=False, train_ds=None, **kwargs: Any) -> List[str]:
    """ike_main.py:210-246.  ``kwargs['sentence_model']`` and ``kwargs['stored_data']`` as in the
    reference; no-eps normalisation of query and bank (fused: the bank is read once, not
    re-normalised into a temporary on every request)."""
    assert train_ds is not None
    sentence_model = kwargs['sentence_model']
    stored_data = kwargs['stored_data']
    new_fact = request['prompt'] + ' ' + request['target']
    query_sentence = _query_sentence(new_fact, new_fact)
    q = _encode_query(sentence_model, query_sentence, device=f'cuda:{hparams.device}')
    bank = bank_from_stored_data(stored_data, device=hparams.device)
    hit = bank.hits(q, hparams.k, normalize_bank="noeps", normalize_query="noeps")[0]
    icl_examples = [stored_data['sentences'][h["corpus_id"]] for h in hit]
    icl_examples.append(query_sentence)
    return icl_examples


def select_icl_samples(prompt, target, stored_data, sentence_model, device, hparams, query_embedding=None
                       ) -> List[str]:
    """multimodal_evaluate.py:395-421.  A supplied ``query_embedding`` is used as given (the
    reference does not re-normalise it); trailing demos are dropped until the joined text is
    at most 2500 characters."""
    new_fact = prompt + ' ' + target
    if query_embedding is not None:
        q = query_embedding.detach().cpu().numpy() if hasattr(query_embedding, "detach") else np.asarray(query_embedding)
        q = np.asarray(q, dtype=np.float32).reshape(-1, np.asarray(q).shape[-1])
        q_norm = None
    else:
        q = _encode_query(sentence_model, _query_sentence(new_fact, new_fact), device=device)
        q_norm = "noeps"
    bank = bank_from_stored_data(stored_data, device=_device_index(device))
    if q_norm is None:
        # bank no-eps normalised, query untouched: normalise the bank side only
        scores, ids = bank.search_host(q, hparams.k, normalize_bank="noeps", normalize_query=None)
        from .bank import hits_from_arrays
        hit = hits_from_arrays(scores, ids)[0]
    else:
        hit = bank.hits(q, hparams.k, normalize_bank="noeps", normalize_query="noeps")[0]
    icl_examples = [stored_data['sentences'][h["corpus_id"]] for h in hit]
    while len(''.join(icl_examples)) > DEMO_CHAR_CAP:
        del icl_examples[-1]
    return icl_examples


# ---- HICE scope gate (multimodal_evaluate.py:382-392, 498-573; HICE.py:238-263) ------------------------------------
_gates: Dict[Tuple[int, int, int], Any] = {}


def _scope_gate(Wo, W_rand, device):
    """Resident ``ScopeGate`` for this (Wo, W_rand) pair: the reference passes both tensors to every call."""
    from .pools import ScopeGate
    dev = _device_index(device)
    for key in [k for k, v in _gates.items() if v[0]() is None or v[1]() is None]:
        _gates.pop(key)
    key = (id(Wo), id(W_rand), dev)
    hit = _gates.get(key)
    versions = (getattr(Wo, "_version", None), getattr(W_rand, "_version", None))
    if hit is not None and hit[0]() is Wo and hit[1]() is W_rand and hit[2] == versions and None not in versions:
        return hit[3]
    gate = ScopeGate(Wo, W_rand, device=dev)
    try:
        _gates[key] = (weakref.ref(Wo), weakref.ref(W_rand), versions, gate)
    except TypeError:
        pass
    return gate


def our_predict_domain(prompt, target, sentence_model, Wo, W_rand, device, mloc_memory=None, threshold=0.81):
    """multimodal_evaluate.py:382-392, same signature and return value ``(logits [n_classes], preds 0-d)``:
    ``relu(q/||q|| @ W_rand) @ Wo.T`` and its argmax, computed by ``scope_gate_kernel``.  Like the reference,
    ``mloc_memory`` / ``threshold`` are accepted and not used here (the memory check lives in ``run_classifier2``)."""
    new_fact = prompt + ' ' + target
    q = _encode_query(sentence_model, _query_sentence(new_fact, new_fact), device=device)
    logits, preds = _scope_gate(Wo, W_rand, device).predict(q)
    return logits[0], preds[0]


def hice_probe_batch(probes: Sequence[Tuple[str, str]], stored_data, sentence_model, device, hparams, Wo, W_rand,
                     mloc_memory=None, threshold: float = 0.81, mem_prompts: Optional[Sequence[str]] = None):
    """The retrieval of ``compute_hice_multimodal_edit_quality`` (multimodal_evaluate.py:498-573) for ALL probes of a
    request -- rephrase, locality, m-loc and the seven dynamic t*i* probes -- in one device call: scope gate
    (``our_predict_domain``) and, for the probes it lets through, ``select_icl_samples``.
    ``probes``: ``(prompt, target)`` pairs.  ``mem_prompts`` (optional): per probe the text of ``run_classifier2``'s
    memory check (``"Prompt: {prompt} \n\n"`` sentence is built here); needs ``mloc_memory``.
    Returns ``[(pred, icl_examples or None), ...]``: ``icl_examples`` as ``select_icl_samples`` returns them (2500-char
    cap applied) where ``pred != 0``, else ``None`` (the reference then evaluates with ``['']``)."""
    sentences = [_query_sentence(p + ' ' + t, p + ' ' + t) for p, t in probes]
    q = np.stack([_encode_query(sentence_model, s, device=device)[0] for s in sentences])
    gate = _scope_gate(Wo, W_rand, device)
    mq = None
    if mloc_memory is not None:
        gate.threshold = float(threshold)
        if gate.memory is None or getattr(gate, "_memory_src", None) is not mloc_memory:
            gate.set_memory(mloc_memory)
            gate._memory_src = mloc_memory
        if mem_prompts is not None:
            mq = np.stack([_encode_query(sentence_model, f"Prompt: {p} \n\n", device=device)[0] for p in mem_prompts])
    else:
        gate.memory = None
    bank = bank_from_stored_data(stored_data, device=_device_index(device))
    _, preds, _, scores, ids = gate.probe(bank, q, hparams.k, mem_query_embeddings=mq)
    preds, ids = preds.cpu().tolist(), ids.cpu().tolist()
    out = []
    for pred, row in zip(preds, ids):
        if pred == 0:
            out.append((0, None))
            continue
        icl = [stored_data['sentences'][i] for i in row if i >= 0]
        while len(''.join(icl)) > DEMO_CHAR_CAP:
            del icl[-1]
        out.append((pred, icl))
    return out
