"""Dataset table of the 20 + 4 benchmark datasets (facts from Chou et al., AAAI 2018, as
listed by the reference's ``baseline.py``): variable count, split sizes, paper PLL and,
where the reference defines them, the four hidden widths ``units`` of the per-variable
auto-encoders.  Same exported name and keys as the reference (``baseline[name]['vars']``,
``['units']``: run.py:41,59).  ``SYNTHETIC_UNITS`` adds the widths chosen for the synthetic
benchmark shapes, which the reference leaves undefined."""

_COLUMNS = ("vars", "train", "valid", "test", "pll", "units")
_ROWS = """
nltcs               16   16181   2157   3236    4.98     15,14,13,12
msnbc               17  291326  38843  58265    6.08     -
kdd                 64  180092  19907  34955    2.07     50,40,30,20
plants              69   17412   2321   3482   10.21     -
audio              100   15000   2000   3000   37.03     80,60,40,30
jester             100    9000   1000   4116   49.75     70,50,40,30
netflix            100   15000   2000   3000   52.67     80,60,40,30
accidents          111   12758   1700   2551   12.69     90,70,50,30
retail             135   22041   2938   4408   10.39     100,70,40,20
pumsb_star         163   12262   1635   2452    9.79     120,90,60,40
dna                180    1600    400   1186   58.46     -
kosarek            190   33375   4450   6675   10.17     140,100,50,25
msweb              294   29441   3270   5000   13.71     -
book               500    8700   1159   1739   35.20     -
tmovie             500    4524   1002    591   58.50     -
webkb              839    2803    558    838  155.51     400,200,100,50
reuters            889    6532   1028   1540   88.55     -
20ng               910   11293   3764   3764  160.82     -
bbc               1058    1670    225    330  256.60     -
ad                1556    2461    327    491    6.01     -
50-17-8            289    5000   2000   2000   49.8696   -
bn2o-30-20-200-2a   50    5000   2000   2000   17.369    -
fs-07             1225    5000   2000   2000   60.0505   -
students_03_02-0000 376   5000   2000   2000    1.4775   -
"""


def _parse():
    table = {}
    for line in _ROWS.strip().splitlines():
        name, nv, tr, va, te, pll, units = line.split()
        row = {"vars": int(nv), "train": int(tr), "valid": int(va), "test": int(te), "pll": float(pll)}
        if units != "-":
            row["units"] = [int(u) for u in units.split(",")]
        table[name] = row
    return table


baseline = _parse()

# widths for the synthetic benchmark shapes (not defined by the reference)
SYNTHETIC_UNITS = {69: [50, 40, 30, 20], 1556: [400, 200, 100, 50]}
