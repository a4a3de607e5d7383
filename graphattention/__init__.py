"""Same import surface as the reference's ``graphattention`` package for the SPUIGACF path
(run_Gowalla.py:23,33): the classes are the B200-native ones from ``ngacf_b200``."""
