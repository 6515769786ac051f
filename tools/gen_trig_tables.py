"""Generates the double-double constants of the correctly-rounded sin/cos used by the yaw branch
(oracle/mpl_oracle.cpp `sincos_cr`, mpl_ros_b200/csrc/mplb_trig.cuh): pi/2 in three parts and 1/n! for n = 2..31.
Run: python tools/gen_trig_tables.py  (needs mpmath; the output is pasted into both files)."""
import mpmath

mpmath.mp.prec = 400


def dd(v):
    hi = float(v)
    lo = float(v - mpmath.mpf(hi))
    return hi, lo


def main():
    p = mpmath.pi / 2
    p1 = float(p)
    p2 = float(p - mpmath.mpf(p1))
    p3 = float(p - mpmath.mpf(p1) - mpmath.mpf(p2))
    print("static const double PIO2_1 = %s, PIO2_2 = %s, PIO2_3 = %s;" % (p1.hex(), p2.hex(), p3.hex()))
    print("static const double TWO_OVER_PI = %s;" % float(2 / mpmath.pi).hex())
    print("/* 1/n! as double-double (hi, lo), n = 2..31 */")
    print("static const double INV_FACT[30][2] = {")
    for n in range(2, 32):
        hi, lo = dd(1 / mpmath.factorial(n))
        print("  {%s, %s}, /* 1/%d! */" % (hi.hex(), lo.hex(), n))
    print("};")


if __name__ == "__main__":
    main()
