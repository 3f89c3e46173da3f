from peclr_b200.rn_25D_wMLPref import RN_25D_wMLPref, ZrootMLP_ref  # noqa: F401
