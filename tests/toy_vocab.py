"""Toy WordPiece vocabulary shared by the reader / tokenizer / driver tests."""
TOY_VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "buffer", "over", "##flow", "in", "the", "parser", ".", ",",
             "sql", "injection", "##s", "crash", "when", "url", "##tag", "is", "null", "a", "b", "fix", "##ed", "!", "use",
             "after", "free", "-", "heap", "cafe", "x", "##y", "##z"]


def write_toy_data(directory):
    """Anchor / CVE / test / validation files in the reference's on-disk formats (reader_memory.py:73-113).
    Shared by tests/test_host.py and oracle/make_reference_golden.py so both read identical bytes."""
    import json
    import os
    anchors = {"CWE-79": "sql injection in the parser", "CWE-120": "buffer overflow", "CWE-416": "use after free"}
    cve = {"CVE-1": {"CWE_ID": "CWE-120", "CVE_Description": "x"}, "CVE-2": {"CWE_ID": "CWE-79", "CVE_Description": "x"},
           "CVE-3": {"CWE_ID": None, "CVE_Description": "x"}}
    rows = [{"Issue_Url": "u0", "Issue_Title": "crash", "Issue_Body": "when url is null", "Security_Issue_Full": 0},
            {"Issue_Url": "u1", "Issue_Title": "buffer overflow", "Issue_Body": "in the parser", "Security_Issue_Full": 1, "CVE_ID": "CVE-1"},
            {"Issue_Url": "u2", "Issue_Title": "fixed", "Issue_Body": "a b", "Security_Issue_Full": "0"},
            {"Issue_Url": "u3", "Issue_Title": "sql", "Issue_Body": "injection", "Security_Issue_Full": "1", "CVE_ID": "CVE-2"},
            {"Issue_Url": "u4", "Issue_Title": "heap", "Issue_Body": "free", "Security_Issue_Full": 1, "CVE_ID": "CVE-3"}]
    paths = {"golden": "CWE_anchor_golden_project.json", "cve": "CVE_dict.json", "test": "test_project.json",
             "validation": "validation_project.json"}
    for key, obj in (("golden", anchors), ("cve", cve), ("test", rows), ("validation", rows)):
        paths[key] = os.path.join(str(directory), paths[key])
        with open(paths[key], "w") as f:
            json.dump(obj, f)
    return paths
