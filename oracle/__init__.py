"""Test infrastructure only: CPU restatement of the reference path and the HF anchor it is pinned to.
Nothing under anyscale_workshop_nyc_2023_b200/ imports this package."""
