"""CPU oracle for the GAT hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  See oracle/pyg_gatconv.py for what is restated and why parity is "unpinned".
"""
