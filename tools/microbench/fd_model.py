# exact model of the FP64 Montgomery product (radix 2^52, 5 limbs, R = 2^260)
import random
p = 21888242871839275222246405745257275088696311157297823662689037894645226208583
M52 = (1<<52)-1
R = 1<<260
n0 = (-pow(p, -1, 1<<52)) % (1<<52)
def trunc53(x):  # round toward zero to 53 significant bits (x integer)
    if x == 0: return 0
    s = -1 if x < 0 else 1
    x = abs(x); b = x.bit_length()
    if b <= 53: return s*x
    sh = b-53
    return s*((x>>sh)<<sh)
def fma_rz(a,b,c): return trunc53(a*b+c)
def bits(x):  # raw IEEE bits of a positive integer-valued double
    assert x > 0
    e = x.bit_length()-1
    assert trunc53(x) == x
    mant = (x << 52 >> e) if e <= 52 else (x >> (e-52))
    assert (mant << e >> 52 if e<=52 else mant << (e-52)) == x
    return ((1023+e) << 52) | (mant & M52)
C1 = 1<<104; C2 = (1<<104)+(1<<52)
def split(a,b):
    hi = fma_rz(a,b,C1); sub = C2-hi; assert trunc53(sub)==sub
    lo = fma_rz(a,b,sub)
    return bits(hi), bits(lo)
def limbs(x): return [(x>>(52*i))&M52 for i in range(5)]
pl = limbs(p)
cnt = lambda k: (min(k,8-k)+1) if 0<=k<=8 else 0
BH = 0x467<<52; BL = 0x433<<52
MASK64=(1<<64)-1
def fd_mul(a,b):
    c = [(-(2*cnt(k)*BL + 2*cnt(k-1)*BH)) & MASK64 for k in range(10)]
    for i in range(5):
        for j in range(5):
            h,l = split(a[i],b[j])
            c[i+j] = (c[i+j]+l)&MASK64; c[i+j+1]=(c[i+j+1]+h)&MASK64
    for k in range(5):
        x = c[k] & M52
        qh, ql = split(x, n0)
        q = ql & M52
        for j in range(5):
            h,l = split(q,pl[j])
            c[k+j] = (c[k+j]+l)&MASK64; c[k+j+1]=(c[k+j+1]+h)&MASK64
        assert c[k] & M52 == 0
        c[k+1] = (c[k+1] + (c[k]>>52)) & MASK64   # c[k] true value: bias fully cancelled by now
        # check c[k] true (small)
        assert c[k] < 1<<60, hex(c[k])
    r=[]
    for k in range(5,10):
        assert c[k] < 1<<60
        r.append(c[k]&M52)
        if k<9: c[k+1] = (c[k+1]+(c[k]>>52))&MASK64
        else: assert c[k]>>52 == 0
    return r
random.seed(1)
for t in range(2000):
    A = random.randrange(0, 16*p); B = random.randrange(0, 2*p)
    if t==0: A=16*p-1; B=2*p-1
    r = fd_mul(limbs(A), limbs(B))
    v = sum(x<<(52*i) for i,x in enumerate(r))
    assert v % p == A*B*pow(R,-1,p) % p
    assert v < 2*p, (v/p)
print("ok", hex(n0))
