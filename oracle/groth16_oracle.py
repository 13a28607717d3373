"""groth16 prove() oracle -- TEST INFRASTRUCTURE (python big ints).

What it restates: `Groth16::prove` (groth16/src/groth16.rs:88-96) = bellman_ce `create_random_proof` on a `CircomCircuit`
(algebraic/src/circom_circuit.rs:94-154: input 0 = ONE, inputs 1..num_inputs-1 public, the rest aux; one `enforce` per
R1CS constraint), with the parameters of `generate_random_parameters` (groth16.rs:82).  bellman_ce 0.3.2 is UN-VENDORED
(Cargo.lock:668-670): the algorithm below is the published Groth16 / bellman one -- PARITY UNPINNED by the reference, whose own
tests only check prove -> verify (groth16.rs:134-262).  What pins this oracle instead:
  * the setup is generated HERE with a known trapdoor (tau, alpha, beta, gamma, delta), so every proof element has a closed form
    in the exponent:  A = alpha + a(tau) + r delta,  B = beta + b(tau) + s delta,
                      C = (sum_aux w_i (beta A_i + alpha B_i + C_i)(tau) + h(tau) Z(tau)) / delta + s A + r B - r s delta
    -- an EXACT expectation for (A, B, C) given (witness, r, s), far stronger than "the verifier accepts";
  * the pairing equation e(A, B) = e(alpha, beta) e(sum_pub w_i IC_i, gamma) e(C, delta) is checked in the exponent
    (bilinearity: A_s B_s = alpha beta + ic_s gamma + C_s delta mod r), which is the verifier's check;
  * file formats: iden3 binary `.r1cs` / `.wtns` as read by algebraic/src/reader.rs:71-138,180-290 (the reference's own
    fixtures test/multiplier.r1cs and groth16/test-vectors/mycircuit_bls12381.r1cs parse), bellman `Parameters::write`
    (vk || h || l || a || b_g1 || b_g2, u32 BE counts, uncompressed big-endian points; the vk part equals the layout that
    the reference's verification_key.bin fixtures have, tests/test_groth16_formats.py).
"""
import struct
from . import curves as C, fr_domain as D

CURVE = {"BN128": dict(g1=C.BN254_G1, g2=C.BN254_G2, field="bn254", nbytes=32),
         "BLS12381": dict(g1=C.BLS381_G1, g2=C.BLS381_G2, field="bls12381", nbytes=48)}


# ---------------------------------------------------------------------------------------------- file formats
def read_r1cs(data):
    """iden3 binary r1cs (algebraic/src/reader.rs:180-290).  Returns dict(prime, n_wires, n_pub_out, n_pub_in, n_prv_in,
    num_inputs, num_aux, constraints=[(A, B, C)] with each lc a list of (wire, coeff))."""
    assert data[:4] == b"r1cs"
    version, n_sections = struct.unpack_from("<II", data, 4)
    assert version == 1
    o = 12; secs = {}
    for _ in range(n_sections):
        t, sz = struct.unpack_from("<IQ", data, o); o += 12
        secs[t] = (o, sz); o += sz
    ho, _ = secs[1]
    fs = struct.unpack_from("<I", data, ho)[0]
    prime = int.from_bytes(data[ho + 4:ho + 4 + fs], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in = struct.unpack_from("<IIII", data, ho + 4 + fs)
    n_labels, n_constraints = struct.unpack_from("<QI", data, ho + 4 + fs + 16)
    co, _ = secs[2]
    cons = []
    for _ in range(n_constraints):
        lcs = []
        for _k in range(3):
            n = struct.unpack_from("<I", data, co)[0]; co += 4
            lc = []
            for _j in range(n):
                w = struct.unpack_from("<I", data, co)[0]; co += 4
                lc.append((w, int.from_bytes(data[co:co + fs], "little"))); co += fs
            lcs.append(lc)
        cons.append(tuple(lcs))
    num_inputs = 1 + n_pub_out + n_pub_in
    return dict(prime=prime, n_wires=n_wires, n_pub_out=n_pub_out, n_pub_in=n_pub_in, n_prv_in=n_prv_in, num_inputs=num_inputs,
                num_aux=n_wires - num_inputs, constraints=cons)


def write_wtns(witness, prime):
    """iden3 `.wtns` (version 2), the layout algebraic/src/reader.rs:87-138 reads"""
    body = b"".join(int(w).to_bytes(32, "little") for w in witness)
    return (b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, 40) + struct.pack("<I", 32) + int(prime).to_bytes(32, "little") + struct.pack("<I", len(witness))
            + struct.pack("<IQ", 2, len(body)) + body)


def read_wtns(data):
    assert data[:4] == b"wtns"
    version, n_sections = struct.unpack_from("<II", data, 4)
    assert version <= 2 and n_sections == 2
    t, sz = struct.unpack_from("<IQ", data, 12); assert t == 1 and sz == 40
    fs = struct.unpack_from("<I", data, 24)[0]; assert fs == 32
    prime = int.from_bytes(data[28:60], "little")
    n = struct.unpack_from("<I", data, 60)[0]
    t, sz = struct.unpack_from("<IQ", data, 64); assert t == 2 and sz == n * fs
    return prime, [int.from_bytes(data[76 + 32 * i:108 + 32 * i], "little") for i in range(n)]


# ---------------------------------------------------------------------------------------------- synthesis (ProvingAssignment)
def synthesize(r1cs, witness, p):
    """bellman `ProvingAssignment`: a, b, c = the three linear combinations of every enforced constraint evaluated on the witness,
    followed by one input-consistency constraint per input (A = input_i, B = C = 0); densities as bellman's DensityTracker keeps them."""
    ni, na = r1cs["num_inputs"], r1cs["num_aux"]
    a, b, c = [], [], []
    a_aux_density = [False] * na; b_input_density = [False] * ni; b_aux_density = [False] * na
    ev = lambda lc: sum(co * witness[w] for w, co in lc) % p
    for A, B, Cc in r1cs["constraints"]:
        if (not A or not B) and not Cc:
            continue                                            # circom_circuit.rs:146
        a.append(ev(A)); b.append(ev(B)); c.append(ev(Cc))
        for w, co in A:
            if co % p and w >= ni: a_aux_density[w - ni] = True
        for w, co in B:
            if co % p:
                if w < ni: b_input_density[w] = True
                else: b_aux_density[w - ni] = True
    for i in range(ni):
        a.append(witness[i] % p); b.append(0); c.append(0)
    return dict(a=a, b=b, c=c, inputs=[w % p for w in witness[:ni]], aux=[w % p for w in witness[ni:]],
                a_aux_density=a_aux_density, b_input_density=b_input_density, b_aux_density=b_aux_density)


def _qap_at_tau(r1cs, field, tau):
    """A_i(tau), B_i(tau), C_i(tau) for every variable and Z(tau), over the domain bellman picks (size m = next power of two >= the
    number of constraints incl. the input-consistency ones)."""
    p = D.MOD[field]
    ni = r1cs["num_inputs"]
    cons = [k for k in r1cs["constraints"] if not ((not k[0] or not k[1]) and not k[2])]
    n_cons = len(cons) + ni
    m = 1
    while m < n_cons: m *= 2
    lg = m.bit_length() - 1
    w = D.omega(field, lg)
    z = (pow(tau, m, p) - 1) % p
    # Lagrange basis at tau: L_j(tau) = z * w^j / (m (tau - w^j))
    minv = pow(m, p - 2, p)
    L = []
    wj = 1
    for _ in range(m):
        L.append(z * wj % p * minv % p * pow((tau - wj) % p, p - 2, p) % p)
        wj = wj * w % p
    nv = r1cs["n_wires"]
    At, Bt, Ct = [0] * nv, [0] * nv, [0] * nv
    for j, (A, B, Cc) in enumerate(cons):
        for wi, co in A: At[wi] = (At[wi] + co * L[j]) % p
        for wi, co in B: Bt[wi] = (Bt[wi] + co * L[j]) % p
        for wi, co in Cc: Ct[wi] = (Ct[wi] + co * L[j]) % p
    for i in range(ni):
        At[i] = (At[i] + L[len(cons) + i]) % p
    return At, Bt, Ct, z, m


def setup(r1cs, curve, trapdoor):
    """`generate_random_parameters` with a KNOWN trapdoor = (tau, alpha, beta, gamma, delta).  Returns the parameters as scalars
    (discrete logs) and as points, in bellman's `Parameters` order."""
    cv = CURVE[curve]; p = D.MOD[cv["field"]]
    g1, g2 = cv["g1"], cv["g2"]
    tau, alpha, beta, gamma, delta = [t % p for t in trapdoor]
    At, Bt, Ct, z, m = _qap_at_tau(r1cs, cv["field"], tau)
    ni = r1cs["num_inputs"]
    ginv, dinv = pow(gamma, p - 2, p), pow(delta, p - 2, p)
    k = [(beta * At[i] + alpha * Bt[i] + Ct[i]) % p for i in range(r1cs["n_wires"])]
    S = dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, tau=tau, z=z, m=m, At=At, Bt=Bt, Ct=Ct,
             ic=[k[i] * ginv % p for i in range(ni)], l=[k[i] * dinv % p for i in range(ni, r1cs["n_wires"])],
             h=[pow(tau, i, p) * z % p * dinv % p for i in range(m - 1)])
    S["a"] = [x for x in At if x]                       # bellman keeps the non-zero points only (and tracks densities in the prover)
    S["b"] = [x for x in Bt if x]
    m1 = lambda s: g1.mul(s, g1.gen) if s else None
    m2 = lambda s: g2.mul(s, g2.gen) if s else None
    P = dict(alpha_g1=m1(alpha), beta_g1=m1(beta), beta_g2=m2(beta), gamma_g2=m2(gamma), delta_g1=m1(delta), delta_g2=m2(delta),
             ic=[m1(s) for s in S["ic"]], h=[m1(s) for s in S["h"]], l=[m1(s) for s in S["l"]], a=[m1(s) for s in S["a"]],
             b_g1=[m1(s) for s in S["b"]], b_g2=[m2(s) for s in S["b"]])
    return S, P


def _enc_fp(v, n): return int(v).to_bytes(n, "big")
def _enc_g1(P, n):
    if P is None: return bytes([0x40]) + bytes(2 * n - 1)
    return _enc_fp(P[0], n) + _enc_fp(P[1], n)
def _enc_g2(P, n):
    if P is None: return bytes([0x40]) + bytes(4 * n - 1)
    return _enc_fp(P[0][1], n) + _enc_fp(P[0][0], n) + _enc_fp(P[1][1], n) + _enc_fp(P[1][0], n)


def write_parameters(P, curve):
    """bellman `Parameters::write`: vk (alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, u32 BE n, ic[n]) then
    h, l, a, b_g1, b_g2, each with a u32 BE count; points uncompressed big-endian (G2: c1 before c0)."""
    n = CURVE[curve]["nbytes"]
    out = _enc_g1(P["alpha_g1"], n) + _enc_g1(P["beta_g1"], n) + _enc_g2(P["beta_g2"], n) + _enc_g2(P["gamma_g2"], n) + _enc_g1(P["delta_g1"], n) + _enc_g2(P["delta_g2"], n)
    out += struct.pack(">I", len(P["ic"])) + b"".join(_enc_g1(q, n) for q in P["ic"])
    for key, enc in (("h", _enc_g1), ("l", _enc_g1), ("a", _enc_g1), ("b_g1", _enc_g1), ("b_g2", _enc_g2)):
        out += struct.pack(">I", len(P[key])) + b"".join(enc(q, n) for q in P[key])
    return out


def prove_in_the_exponent(r1cs, curve, S, witness, r, s):
    """(A_s, B_s, C_s): the discrete logs of the proof bellman's create_proof produces for (witness, r, s); plus the public-input
    combination ic_s.  Everything mod the scalar field."""
    cv = CURVE[curve]; p = D.MOD[cv["field"]]
    ni = r1cs["num_inputs"]
    w = [x % p for x in witness]
    a_t = sum(w[i] * S["At"][i] for i in range(len(w))) % p
    b_t = sum(w[i] * S["Bt"][i] for i in range(len(w))) % p
    c_t = sum(w[i] * S["Ct"][i] for i in range(len(w))) % p
    h_t = (a_t * b_t - c_t) * pow(S["z"], p - 2, p) % p            # h(tau); A B - C vanishes on the domain iff the witness satisfies the R1CS
    dinv = pow(S["delta"], p - 2, p)
    A_s = (S["alpha"] + a_t + r * S["delta"]) % p
    B_s = (S["beta"] + b_t + s * S["delta"]) % p
    l_t = sum(w[ni + i] * S["l"][i] for i in range(len(w) - ni)) % p
    C_s = (l_t + h_t * S["z"] % p * dinv + s * A_s + r * B_s - r * s % p * S["delta"]) % p
    ic_s = sum(w[i] * S["ic"][i] for i in range(ni)) % p
    return A_s, B_s, C_s, ic_s


def pairing_equation_holds(S, p, A_s, B_s, C_s, ic_s):
    """e(A, B) = e(alpha, beta) e(IC(pub), gamma) e(C, delta), in the exponent"""
    return A_s * B_s % p == (S["alpha"] * S["beta"] + ic_s * S["gamma"] + C_s * S["delta"]) % p


def h_coefficients(r1cs, curve, witness):
    """the quotient's coefficients as bellman computes them (ifft, coset fft, pointwise, divide by Z on the coset, icoset fft):
    used to check the GPU's H against the definition (A B - C = H Z) independently of the trapdoor."""
    cv = CURVE[curve]; field = cv["field"]; p = D.MOD[field]
    syn = synthesize(r1cs, witness, p)
    m = 1
    while m < len(syn["a"]): m *= 2
    pad = lambda v: v + [0] * (m - len(v))
    return D.groth16_h(field, pad(syn["a"]), pad(syn["b"]), pad(syn["c"])), m
