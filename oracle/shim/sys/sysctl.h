/* Empty stand-in: glibc >= 2.32 dropped <sys/sysctl.h>, which the reference's
 * core/parallel.cpp:46 still includes on Linux (it only uses it on BSD/macOS).
 * Test infrastructure for building oracle/_ref; not part of the product. */
