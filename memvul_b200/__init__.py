"""memvul_b200 -- B200-native batch inference for MemVul (BERT encoder + CWE-anchor memory match)."""
__version__ = "0.1.0"
