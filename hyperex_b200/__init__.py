"""hyperex_b200 — B200-native primer-search-and-extract hot path of HyperEx (C ABI + host mirror)."""
from .api import (ALPHABET_NONE, FWD_ANY, PAIR_VALID, RESULT_DTYPE, REV_ANY, RNA, SUPPORTED_REGIONS, Alphabet,
                  Context, HyperexError, bind_to_gpu_numa_node, default_primers, file_to_vec, int_peak, get_hypervar_regions, pinned_empty, plan_groups, plan_peq_word, plan_packing,
                  pinned_free, primers_to_region, pack_records, PACK_CODES, REC_HAS_T, REC_HAS_U, REC_NON_IUPAC, read_fasta, read_fasta_packed, region_to_primer, sequence_type, supported_regions, to_complement,
                  to_reverse_complement)

__all__ = [n for n in dir() if not n.startswith("_")]
