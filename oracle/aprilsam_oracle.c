/* aprilsam_oracle.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 * Placeholder for the plain-C restatement of the reference's Gauss-Newton path.  This round the
 * parity oracle is the UNMODIFIED reference itself, compiled from /root/reference by
 * oracle/Makefile into oracle/_ref/ (deterministic clock, see oracle/ref_clock.c) and pinned by
 * the golden vectors under tests/golden/ (tools/make_golden.py; known answers of SURVEY.md
 * section 8c are asserted in tests/test_host_cpu.py::test_golden_known_answers).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may use
 * anything under oracle/.
 */
int aprilsam_oracle_port_available(void) { return 0; }
