"""CPU oracle for the DE-VQA fact-bank retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``de-vqa_b200/`` may import this
package.  The only permitted importers are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only
as the checker or the timed CPU baseline, never as the product path.

PARITY UNPINNED (arithmetic): the reference keeps the arithmetic of this path in
the un-vendored, un-pinned PyPI package ``sentence-transformers``
(``sentence_transformers.util.{normalize_embeddings,dot_score,semantic_search}``),
which is neither under /root/reference nor installed in this image, and the
reference ships no tests, golden vectors or ``.pkl`` banks for it (SURVEY.md
§4, §8c).  ``oracle/semantic_search.py`` therefore restates the *published*
algorithm of those helpers on torch-CPU (the same ``F.normalize`` / ``torch.mm``
/ ``torch.topk`` / ``heapq`` primitives the upstream bottoms out in).

PINNED (glue): the reference-owned host logic around the search --
``finds_sim`` (dataset/vllm.py:65-87), ``select_icl_samples``
(easyeditor/evaluate/multimodal_evaluate.py:395-421),
``apply_hice_to_multimodal_model`` (easyeditor/models/ike/ike_main.py:210-246),
``encode_ike_facts_multimodal`` (easyeditor/models/ike/util.py:54-86) -- is
pinned by fixtures under ``tests/golden/`` that were produced by executing the
reference's own function bodies in this container
(``tests/golden/make_golden.py``), with ``sentence_transformers`` provided by
``oracle/st_shim.py``.
"""
