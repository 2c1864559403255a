"""CPU oracle for the OA-DG hot path (OA-Mix transform + OA-Loss contrastive loss).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or the timed CPU baseline.  The product (``oadg_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Contents
--------
``saliency_np``   restatement of opencv-contrib ``StaticSaliencySpectralResidual``
                  (absent from this image; PARITY UNPINNED for that one piece).
``prims_np``      integer / float restatements of the third-party primitives the
                  reference calls (cv2.warpAffine 8U, Pillow ImageOps LUTs,
                  blurred-mask profiles), validated against the live cv2/Pillow.
``oamix_np``      restatement of ``OAMix`` (reference ``oa_mix.py:32-313`` and
                  ``bbox_augmentation.py`` / ``augmix.py`` call sites), pinned
                  bit-exactly against the reference run under ``ref_loader``.
``supcon_np``     restatement of ``supcontrast`` / ``ContrastiveLossPlus``
                  (``contrastive_loss.py:147-232``, ``contrastive_loss_plus.py:31-50``).
``ref_loader``    imports the UNMODIFIED reference modules from /root/reference
                  under an mmcv stub (dev container only; never on the GPU box).
"""
