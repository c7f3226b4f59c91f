"""Label tables of the PileupModel heads (contract shared with PileupModel/options.py:3-30 of the reference)."""
base_idx = {"A": 0, "C": 1, "G": 2, "T": 3}
gt_decoded_labels = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT",
                     "DD", "AD", "CD", "GD", "TD", "II", "AI", "CI", "GI", "TI", "ID"]
zy_decoded_labels = ["0/0", "1/1", "0/1"]
