"""CPU oracle for the ThinkDiff aligner hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or as the
reported CPU baseline -- never as the thing shipped.  The product (``thinkdiff_mlre_b200``) fails loudly when its
CUDA library is missing; it never falls back to this code.

Contents
--------
``aligner_ref``  restatement of ``build_vision_projector`` (``mlp2x_gelu_t5_norm``) forward/backward
                 (reference: thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:41-79, T5LayerNorm from
                 transformers==4.46.1 modeling_t5.py).
``pack_ref``     restatement of the ragged pad/stack/mask collater
                 (reference: thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py:34-185).
``loss_ref``     CrossEntropyLoss(ignore_index=-100) as used at ...embed_decoder_2.py:241-246, and the
                 masked MSE that BASELINE.json's north_star adds (not in the reference).
``t5_head_ref``  frozen T5 output head + CE and the frozen bias-free Linear (...embed_decoder_2.py:231-246), f-1 first slice.
``t5_decoder_ref`` the frozen T5 v1.1 decoder stack (third-party: transformers==4.46.1 modeling_t5.py) restated on PACKED
                 aligner rows -- varlen cross-attention, relative position bias, gated GELU; pinned live against the installed
                 transformers in tests/test_t5_decoder_oracle.py. No kernel consumes it yet (f-1 groundwork).
``ref_loader``   ast-extracts and exec's the reference's OWN functions from /root/reference (dev container only).
``make_golden``  regenerates tests/golden/*.npz by running the reference's own code (dev container only).

Parity pinning: the reference ships no tests, fixtures or golden vectors for this path, so the
oracle is pinned against the reference's own code executed here: ``tests/test_oracle_vs_reference.py`` runs live
when /root/reference exists, and the committed ``tests/golden`` vectors (made by ``make_golden.py`` from the exec'd
reference functions) pin it everywhere else, including the GPU box.
"""
