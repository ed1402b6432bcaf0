"""CPU oracle for the RAW->sRGB->bitstream hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it,
and only as the checker.  The product (``realcamnet_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Contents
--------
``cai.py``        restatement of the CompressAI subset the reference calls
                  (un-vendored third-party dependency; parity UNPINNED vs upstream,
                  see header of that file)
``rans.py``       pure-Python/numpy rANS64 coder + pmf->quantized-CDF (restated)
``rans_c.c``      the same coder in plain C (fast checker for multi-million symbol cases)
``refpath.py``    functional torch-CPU fp32 restatement of the reference forward path
                  (raw_compression_tcm_final, LiteISPNet_GFM_LSC, GMA_Block, ...)
``ref_import.py`` imports the *unmodified* reference from /root/reference (only in
                  the authoring container) behind stub modules, to validate the
                  restatement and to generate tests/golden fixtures
``weights.py`` / ``inputs.py``  re-exports of realcamnet_b200/synthetic.py (seeded inputs and name-keyed deterministic
                  weights live outside this package so that bench.py and the tools never import it)
"""
