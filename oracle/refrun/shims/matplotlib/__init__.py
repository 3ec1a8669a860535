"""Call-swallowing ``matplotlib`` stub (plot_results.py:2-3 imports it at
module scope).  TEST INFRASTRUCTURE ONLY."""
