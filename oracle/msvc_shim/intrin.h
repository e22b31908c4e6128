/* Stand-in for MSVC's <intrin.h>; see ref_tree_msvc.h.  TEST INFRASTRUCTURE ONLY. */
