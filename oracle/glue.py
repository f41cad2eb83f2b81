"""CPU restatement of the reference-owned host logic around the search.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function cites the
reference lines it follows; ``tests/golden/make_golden.py`` pins them against
the reference's own function bodies executed in this container.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import semantic_search as ss

DEMO_CHAR_CAP = 2500   # multimodal_evaluate.py:418


def fact_sentences(prompt: str, target: str, rephrase: str, loc_prompt: str, loc_answer: str):
    """The three bank sentences one record contributes (util.py:65,71-73)."""
    fact = prompt + ' ' + target
    return [
        f"New Fact: {fact}\nPrompt: {fact}\n\n",
        f"New Fact: {fact}\nPrompt: {rephrase} {target}\n\n",
        f"New Fact: {fact}\nPrompt: {loc_prompt} {loc_answer}\n\n",
    ]


def build_fact_bank(records: Sequence[Dict], encoder) -> Dict:
    """Bank dict in the frozen pickle layout (util.py:54-86).

    Row order is (edit, rephrase, locality) interleaved per record; keys are
    ``sentences``, ``embeddings`` (np.float32 [3n, D]), ``images``, ``prompts``.
    """
    sentences: List[str] = []
    images: List[str] = []
    prompts: List[List[str]] = []
    for rec in records:
        images += [rec['image_path'], rec['rephrase_image_path'], rec['locality_image_path']]
        prompts += [[rec['prompt'], rec['target']],
                    [rec['rephrase_prompt'], rec['target']],
                    [rec['locality_prompt'], rec['locality_ground_truth']]]
        sentences += fact_sentences(rec['prompt'], rec['target'], rec['rephrase_prompt'],
                                    rec['locality_prompt'], rec['locality_ground_truth'])
    return {'sentences': sentences, 'embeddings': encoder.encode(sentences),
            'images': images, 'prompts': prompts}


def query_sentence(prompt: str, target: str) -> str:
    """Query template of the multimodal IKE/HICE lookups (ike_main.py:196-198,224-228;
    multimodal_evaluate.py:396-398).  ``apply_ike_to_model`` (text) uses
    ``Prompt: {prompt}`` instead (ike_main.py:39-40)."""
    fact = prompt + ' ' + target
    return f"New Fact: {fact}\nPrompt: {fact}\n\n"


def finds_sim(encoder, bank_norm: torch.Tensor, prompts, images, src: str, trg: str,
              tops: int = 5) -> Tuple[List[str], str]:
    """dataset/vllm.py:65-87: top-``tops`` search, first hit whose stored target
    differs from ``trg``, else the last hit.  ``bank_norm`` is already normalised
    (dataset/vllm.py:104)."""
    q = ss.normalize_embeddings(torch.tensor(encoder.encode(src, show_progress_bar=False)).unsqueeze(0))
    hit = ss.semantic_search(q, bank_norm, score_function=ss.dot_score, top_k=tops)[0]
    for h in hit:
        if prompts[h["corpus_id"]][1] != trg:
            return prompts[h["corpus_id"]], images[h["corpus_id"]]
    return prompts[hit[-1]["corpus_id"]], images[hit[-1]["corpus_id"]]


def ike_demos(encoder, bank_raw, sentences, prompt: str, target: str, k: int) -> List[str]:
    """apply_ike_to_multimodal_model (ike_main.py:193-206): eps-normalise bank and
    query, top-k, demos = bank sentences + the new fact itself."""
    bank = ss.normalize_embeddings(torch.as_tensor(np.asarray(bank_raw)))
    q = ss.normalize_embeddings(torch.tensor(
        encoder.encode(query_sentence(prompt, target), show_progress_bar=False)).unsqueeze(0))
    hit = ss.semantic_search(q, bank, score_function=ss.dot_score, top_k=k)[0]
    return [sentences[h["corpus_id"]] for h in hit] + [query_sentence(prompt, target)]


def hice_demos(encoder, bank_raw: torch.Tensor, sentences, prompt: str, target: str, k: int) -> List[str]:
    """apply_hice_to_multimodal_model (ike_main.py:224-246): no-eps normalisation
    of query and of the whole bank per call."""
    q = torch.tensor(encoder.encode(query_sentence(prompt, target), show_progress_bar=False)).unsqueeze(0)
    q = q / torch.norm(q, p=2, dim=-1)
    bank = ss.normalize_noeps(bank_raw)
    hit = ss.semantic_search(q, bank, score_function=ss.dot_score, top_k=k)[0]
    return [sentences[h["corpus_id"]] for h in hit] + [query_sentence(prompt, target)]


def select_icl_samples(encoder, bank_raw: torch.Tensor, sentences, prompt: str, target: str,
                       k: int, query_embedding: torch.Tensor = None) -> List[str]:
    """multimodal_evaluate.py:395-421: like ``hice_demos`` without the trailing new
    fact, with trailing demos dropped until the joined text is <= 2500 chars; a
    caller-supplied ``query_embedding`` is used as-is (not re-normalised)."""
    if query_embedding is None:
        q = torch.tensor(encoder.encode(query_sentence(prompt, target), show_progress_bar=False)).unsqueeze(0)
        q = q / torch.norm(q, p=2, dim=-1)
    else:
        q = query_embedding
    bank = ss.normalize_noeps(torch.as_tensor(bank_raw))
    hit = ss.semantic_search(q, bank, score_function=ss.dot_score, top_k=k)[0]
    demos = [sentences[h["corpus_id"]] for h in hit]
    while len(''.join(demos)) > DEMO_CHAR_CAP:
        del demos[-1]
    return demos
