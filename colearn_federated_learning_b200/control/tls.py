"""Optional TLS for the two TCP planes (MQTT bus and worker RPC).

The reference leaves TLS as two commented-out keys on the PySyft websocket server (``cert_path`` / ``key_path``,
``remote_worker.py:60-61``).  Here both planes take a standard :class:`ssl.SSLContext`:

* servers (``TcpBroker``, ``WorkerServer``): :func:`server_context` — certificate + key, optionally a CA bundle with
  ``require_client_cert=True`` for mutual TLS (only devices holding a certificate signed by the deployment's CA can
  announce themselves or be asked to train);
* clients (``BusClient.tls_set`` like paho, ``RemoteWorkerClient``): :func:`client_context`.

:func:`make_test_pki` writes a throw-away CA + server/client certificates (used by the tests and handy for a lab
set-up); it needs the ``cryptography`` package, nothing else here does.
"""
from __future__ import annotations

import datetime
import ipaddress
import os
import ssl
from typing import Dict, Optional


def server_context(certfile: str, keyfile: str, cafile: Optional[str] = None, require_client_cert: bool = False) -> ssl.SSLContext:
    ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
    ctx.minimum_version = ssl.TLSVersion.TLSv1_2
    ctx.load_cert_chain(certfile, keyfile)
    if cafile:
        ctx.load_verify_locations(cafile)
    if require_client_cert:
        if not cafile:
            raise ValueError("require_client_cert needs the CA bundle that signed the client certificates")
        ctx.verify_mode = ssl.CERT_REQUIRED
    return ctx


def client_context(cafile: Optional[str] = None, certfile: Optional[str] = None, keyfile: Optional[str] = None,
                   check_hostname: bool = True) -> ssl.SSLContext:
    """``cafile=None`` trusts the system store; devices are addressed by IP, so certificates need IP SANs (or pass
    ``check_hostname=False`` to verify the chain only)."""
    ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_CLIENT)
    ctx.minimum_version = ssl.TLSVersion.TLSv1_2
    if cafile:
        ctx.load_verify_locations(cafile)
    else:
        ctx.load_default_certs()
    ctx.check_hostname = check_hostname
    if certfile:
        ctx.load_cert_chain(certfile, keyfile)
    return ctx


def contexts_from_cli(cafile: Optional[str], certfile: Optional[str], keyfile: Optional[str], server: bool,
                      require_client_cert: bool = False, check_hostname: bool = True) -> Optional[ssl.SSLContext]:
    """CLI helper: no TLS flags → None (plain TCP, the reference's behaviour).  Client contexts verify the peer's name /
    IP SAN by default — with a deployment-wide CA, chain verification alone would let any device holding *a* certificate
    pose as the broker or as another device; ``--tls-no-verify-hostname`` is the explicit opt-out."""
    if not (cafile or certfile):
        return None
    if server:
        if not (certfile and keyfile):
            raise SystemExit("a TLS server needs --tls-cert and --tls-key")
        return server_context(certfile, keyfile, cafile, require_client_cert)
    return client_context(cafile, certfile, keyfile, check_hostname=check_hostname)


def make_test_pki(directory: str, hosts=("127.0.0.1", "localhost")) -> Dict[str, str]:
    """Write ``ca.pem``, ``server.pem/.key``, ``client.pem/.key`` (EC P-256, valid 30 days) into ``directory``."""
    from cryptography import x509
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.x509.oid import NameOID

    os.makedirs(directory, exist_ok=True)
    now = datetime.datetime.now(datetime.timezone.utc)

    def name(cn):
        return x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, cn)])

    def write(path, data):
        with open(path, "wb") as f:
            f.write(data)
        return path

    def key_pem(key):
        return key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.PKCS8, serialization.NoEncryption())

    ca_key = ec.generate_private_key(ec.SECP256R1())
    ca_cert = (x509.CertificateBuilder().subject_name(name("colearn test CA")).issuer_name(name("colearn test CA"))
               .public_key(ca_key.public_key()).serial_number(x509.random_serial_number())
               .not_valid_before(now - datetime.timedelta(minutes=5)).not_valid_after(now + datetime.timedelta(days=30))
               .add_extension(x509.BasicConstraints(ca=True, path_length=None), critical=True)
               .sign(ca_key, hashes.SHA256()))
    out = {"ca": write(os.path.join(directory, "ca.pem"), ca_cert.public_bytes(serialization.Encoding.PEM))}
    sans = []
    for h in hosts:
        try:
            sans.append(x509.IPAddress(ipaddress.ip_address(h)))
        except ValueError:
            sans.append(x509.DNSName(h))
    for role in ("server", "client"):
        key = ec.generate_private_key(ec.SECP256R1())
        builder = (x509.CertificateBuilder().subject_name(name(f"colearn {role}")).issuer_name(ca_cert.subject)
                   .public_key(key.public_key()).serial_number(x509.random_serial_number())
                   .not_valid_before(now - datetime.timedelta(minutes=5)).not_valid_after(now + datetime.timedelta(days=30))
                   .add_extension(x509.SubjectAlternativeName(sans), critical=False))
        cert = builder.sign(ca_key, hashes.SHA256())
        out[role + "_cert"] = write(os.path.join(directory, role + ".pem"), cert.public_bytes(serialization.Encoding.PEM))
        out[role + "_key"] = write(os.path.join(directory, role + ".key"), key_pem(key))
    return out
